// bf16 GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma issued by one elected thread, operands
// staged in shared memory by TMA (cp.async.bulk.tensor, 128B swizzle), fp32 accumulators in TMEM, read back with
// tcgen05.ld for a fused epilogue.
//
//   C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N])        A, W bf16 K-major; fp32 accumulate
//
// Used for every dense contraction of the caption path in GIC_DTYPE_BF16 / BF16X2: Conv1D c_attn / c_proj / c_fc /
// mlp.c_proj (HF:models/gpt2/modeling_gpt2.py:185,223,238-243), the mapping-network Linears (src/models.py:52-56,119,
// 129-139) and the tied LM head (HF :705-706) whose epilogue is fused with the greedy argmax (src/models.py:398-443):
// each CTA reduces its 128 x BLOCK_N logit tile to one (max, lowest index) pair per row, so logits never reach HBM.
//
// Kernel shape: PERSISTENT, one CTA per SM walking 128 x BLOCK_N output tiles; BLOCK_K = 64 (one 128-byte swizzle atom),
// a 6..10-stage TMA->MMA mbarrier ring that keeps streaming across tile boundaries, and two TMEM accumulators so the
// epilogue of tile i overlaps the main loop of tile i+1.  6 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer, warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Tiles are numbered M-fastest: the ~148 tiles in
// flight share a handful of W tiles, each fetched from HBM once and served from L2 to the other M tiles.
// BF16X2 ("split") mode: A = A_hi + A_lo, W = W_hi + W_lo (each bf16); three MMAs per k-step
// (hi.hi + hi.lo + lo.hi) into the same accumulator give ~16 mantissa bits.
#include <cuda.h>

#include "kernels.cuh"

namespace gic {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace ptx

// ---------------------------------------------------------------------------------------------------------------
// Descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ---------------------------------------------------------------------------------------------------------------
// K-major operand tile in smem written by TMA with 128B swizzle: rows of 128 bytes, 8-row swizzle atoms (1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units          bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major) bits [16,30)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset: 8 rows x 128 B    bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) /* c_format = F32 */ | (1u << 7) /* a_format = BF16 */ | (1u << 10) /* b_format = BF16 */ |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------------------------
struct alignas(64) GemmKernelParams {
  TmaDesc a_hi, w_hi, a_lo, w_lo;
  int M, N, K;
  int epilogue;
  int mma_repeat;  // 1; > 1 only in the microbenchmark probe: re-issue each k-block's MMAs to measure the tensor-pipe rate
  const float* bias;
  float* out_f32; int ld_f32;
  bf16* out_hi; bf16* out_lo; int ld_bf16;
  float* part_val; int* part_idx;
  long long* trace;  // null; microbenchmark only: clock64 timeline of CTA 0's producer / MMA / epilogue threads
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 128;  // two 128-byte swizzle atoms per stage, each operand fetched by ONE 3-D TMA
constexpr int GEMM_ATOM_K = 64;    // elements per swizzle atom row
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N, bool SPLIT>
struct GemmTile {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // [2 atoms][128 rows][128 B]
  static constexpr int W_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;       // [2 atoms][BLOCK_N rows][128 B]
  static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (SPLIT ? 2 : 1);
  // one persistent CTA per SM: spend (almost) all of its shared memory on the TMA ring.  Measured round 1
  // (microbench tma_probe): the single-thread producer loop costs ~520 cycles per stage (mbarrier wait + expect_tx + two
  // TMA issues) whatever the stage size -- a 64-wide k-block capped the feed at 123 GB/s per SM and the GEMMs ran at ~1050
  // cycles per k-block, 4x the MMA time; 128-wide k-blocks fetched by one 3-D TMA per operand halve the per-byte overhead.
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static_assert(STAGES * STAGE_BYTES + 4096 <= 227 * 1024, "tile does not fit in shared memory");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 512 /* barriers */ + 2 * BLOCK_N * 4 /* bias */;
  static constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;   // one accumulator buffer
  static constexpr int TMEM_NEED = 2 * ACC_COLS;                 // double-buffered: epilogue(i) overlaps main loop(i+1)
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;  // power of two
};

// Persistent kernel: grid = min(#tiles, #SMs); CTA c walks tiles c, c + grid, ...  (tile t -> m_tile = t % m_tiles,
// n_tile = t / m_tiles, so the ~148 tiles in flight share few W tiles and each is fetched from HBM once).
template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmKernelParams p) {
  using Tile = GemmTile<BLOCK_N, SPLIT>;
  constexpr int STAGES = Tile::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Tile::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Tile::STAGE_BYTES + 512);  // [2][BLOCK_N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  const int nk = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  pdl_launch_dependents();  // let the next kernel's CTAs be scheduled behind this grid (see common.cuh)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  const long long t_start = tracing ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.a_hi);
    ptx::prefetch_tmap(&p.w_hi);
    if (SPLIT) {
      ptx::prefetch_tmap(&p.a_lo);
      ptx::prefetch_tmap(&p.w_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 4);  // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_base_slot, Tile::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (tracing && threadIdx.x == 0) p.trace[600] = clock64() - t_start;  // prologue done
  pdl_wait();  // prologue above overlapped the previous kernel; its outputs are visible from here on

  if (warp == 0) {
    // ===== TMA producer: streams k-blocks of successive tiles through the ring without pausing at tile boundaries =====
    if (lane == 0) {
      uint32_t it = 0;  // k-block counter across all tiles of this CTA
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile % m_tiles) * GEMM_BLOCK_M, n0 = (tile / m_tiles) * BLOCK_N;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (tracing && it < 60) p.trace[it * 4 + 0] = clock64() - t_start;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          if (tracing && it < 60) p.trace[it * 4 + 1] = clock64() - t_start;
          uint8_t* st = smem + s * Tile::STAGE_BYTES;
          ptx::mbar_expect_tx(&full_bar[s], Tile::STAGE_BYTES);
          const int ka = kb * (GEMM_BLOCK_K / GEMM_ATOM_K);  // first swizzle atom of this k-block (3rd tensor-map coordinate)
          ptx::tma_load_3d(st, &p.a_hi, &full_bar[s], 0, m0, ka);
          ptx::tma_load_3d(st + Tile::A_BYTES, &p.w_hi, &full_bar[s], 0, n0, ka);
          if (SPLIT) {
            ptx::tma_load_3d(st + Tile::A_BYTES + Tile::W_BYTES, &p.a_lo, &full_bar[s], 0, m0, ka);
            ptx::tma_load_3d(st + 2 * Tile::A_BYTES + Tile::W_BYTES, &p.w_lo, &full_bar[s], 0, n0, ka);
          }
          if (tracing && it < 60) p.trace[it * 4 + 2] = clock64() - t_start;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread), alternating between the two TMEM accumulators =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 1, use = local >> 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], (use & 1) ^ 1);  // epilogue has drained this accumulator (passes at first use)
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * Tile::ACC_COLS;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (tracing && it < 60) p.trace[256 + it * 4 + 0] = clock64() - t_start;
          ptx::mbar_wait(&full_bar[s], ph);
          if (tracing && it < 60) p.trace[256 + it * 4 + 1] = clock64() - t_start;
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * Tile::STAGE_BYTES);
          const uint64_t a_hi = make_smem_desc_sw128(sa);
          const uint64_t w_hi = make_smem_desc_sw128(sa + Tile::A_BYTES);
          const uint64_t a_lo = make_smem_desc_sw128(sa + Tile::A_BYTES + Tile::W_BYTES);
          const uint64_t w_lo = make_smem_desc_sw128(sa + 2 * Tile::A_BYTES + Tile::W_BYTES);
          for (int rep = 0; rep < p.mma_repeat; ++rep)
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // k-step inside the stage: atom (k / 4) is a whole [rows][128 B] tile further on, then 32 bytes per step inside it
            constexpr uint64_t A_ATOM = (uint64_t)(GEMM_BLOCK_M * 128) >> 4, W_ATOM = (uint64_t)(BLOCK_N * 128) >> 4;
            const uint64_t ka = (uint64_t)(k / 4), ki = (uint64_t)(((k % 4) * 16 * 2) >> 4);
            const uint64_t aoff = ka * A_ATOM + ki, woff = ka * W_ATOM + ki;
            ptx::umma_bf16(tmem_acc, a_hi + aoff, w_hi + woff, idesc, (kb | k | rep) != 0);
            if (SPLIT) {
              ptx::umma_bf16(tmem_acc, a_hi + aoff, w_lo + woff, idesc, 1);
              ptx::umma_bf16(tmem_acc, a_lo + aoff, w_hi + woff, idesc, 1);
            }
          }
          ptx::umma_commit(&empty_bar[s]);  // smem slot is free once these MMAs have read it
          if (tracing && it < 60) p.trace[256 + it * 4 + 2] = clock64() - t_start;
        }
        ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    // Everything that does not depend on the accumulator is fetched BEFORE waiting for it (bias tile -> smem, first
    // residual chunk -> registers), and inside the loop the next chunk's residual is loaded before the current chunk is
    // stored: output and residual alias (in-place +=), so loads placed after stores would serialise one L2 round trip per
    // float4 (ncu source page, round 1).
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = (warp - 2) * 32 + lane;  // 0..127 over the four epilogue warps
    const bool resid = p.epilogue == EPI_RESIDUAL;
    const bool vec_f32 = (p.ld_f32 % 4 == 0);
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const uint32_t acc = local & 1, use = local >> 1;
      const int n_tile = tile / m_tiles;
      const int m0 = (tile % m_tiles) * GEMM_BLOCK_M, n0 = n_tile * BLOCK_N;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float* bias_s = s_bias + acc * BLOCK_N;
      for (int c = et; c < BLOCK_N; c += 128) bias_s[c] = (p.bias && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      float4 res_next[8];
      auto load_res = [&](int c0) {
        const float* src = p.out_f32 + (size_t)row * p.ld_f32 + n0 + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) res_next[j] = *reinterpret_cast<const float4*>(src + 4 * j);
      };
      const bool res_vec_ok = resid && row_ok && vec_f32;
      if (res_vec_ok && n0 + 32 <= p.N) load_res(0);
      if (tracing && et == 0 && local < 8) p.trace[512 + local * 4 + 0] = clock64() - t_start;
      ptx::mbar_wait(&tmem_full_bar[acc], use & 1);
      if (tracing && et == 0 && local < 8) p.trace[512 + local * 4 + 1] = clock64() - t_start;
      ptx::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * Tile::ACC_COLS;
      float best = -INFINITY;
      int best_idx = 0x7fffffff;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        if (c0 + 32 >= BLOCK_N) {
          // last TMEM read of this tile: hand the accumulator back to the MMA warp before doing the math / stores
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int col0 = n0 + c0;
        if (col0 >= p.N) continue;  // warp-uniform
        const bool full = (col0 + 32 <= p.N);
        float4 res_cur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) res_cur[j] = res_next[j];
        if (res_vec_ok && c0 + 32 < BLOCK_N && col0 + 64 <= p.N) load_res(c0 + 32);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]) + bias_s[c0 + j];
          if (p.epilogue == EPI_TANH) x = tanhf(x);
          else if (p.epilogue == EPI_GELU) x = SPLIT ? gelu_tanh(x) : gelu_tanh_fast(x);
          else if (p.epilogue == EPI_RELU) x = fmaxf(x, 0.f);
          v[j] = x;
        }
        if (p.part_val) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < p.N && v[j] > best) {  // strict > keeps the lowest index among equal maxima
              best = v[j];
              best_idx = col;
            }
          }
        }
        if (!row_ok) continue;
        if (p.out_f32) {
          float* dst = p.out_f32 + (size_t)row * p.ld_f32 + col0;
          if (full && vec_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (resid) { o.x += res_cur[j].x; o.y += res_cur[j].y; o.z += res_cur[j].z; o.w += res_cur[j].w; }
              *reinterpret_cast<float4*>(dst + 4 * j) = o;
            }
          } else {
            float old[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = (resid && col0 + j < p.N) ? dst[j] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) dst[j] = old[j] + v[j];
          }
        }
        if (p.out_hi) {
          bf16* dh = p.out_hi + (size_t)row * p.ld_bf16 + col0;
          bf16* dl = p.out_lo ? p.out_lo + (size_t)row * p.ld_bf16 + col0 : nullptr;
          if (full && (p.ld_bf16 % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              Vec16<bf16> hv;
              hv.pack(v + j);
              hv.store(dh + j);
              if (dl) {
                float hf[8], lf[8];
                hv.unpack(hf);
#pragma unroll
                for (int t = 0; t < 8; ++t) lf[t] = v[j + t] - hf[t];
                Vec16<bf16> lv;
                lv.pack(lf);
                lv.store(dl + j);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) {
                const bf16 hb = __float2bfloat16_rn(v[j]);
                dh[j] = hb;
                if (dl) dl[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hb));
              }
          }
        }
      }
      if (p.part_val && row_ok) {
        p.part_val[(size_t)n_tile * p.M + row] = best;
        p.part_idx[(size_t)n_tile * p.M + row] = best_idx;
      }
      if (tracing && et == 0 && local < 8) p.trace[512 + local * 4 + 2] = clock64() - t_start;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Tile::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int tma_init() {
  if (g_encode) return GIC_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  GIC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  GIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from this driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return GIC_OK;
}

// bf16 row-major [rows, cols] viewed as [cols/64 atoms][rows][64] so that ONE box of (64 x box_rows x 2 atoms) lands in
// shared memory as two consecutive 128B-swizzled K-major tiles -- exactly the layout the UMMA descriptors walk.
// Out-of-bounds rows / atoms are zero-filled.  Requires cols % 8 == 0; a partial last atom is zero-filled by the box bounds.
int make_tma_2d_bf16(TmaDesc* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
  GIC_TRY(tma_init());
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
  GIC_REQUIRE(((uintptr_t)base % 16) == 0 && (row_stride_elems * 2) % 16 == 0,
              "TMA: base must be 16-byte aligned and the row stride a multiple of 8 elements (stride %llu)",
              (unsigned long long)row_stride_elems);
  GIC_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA: box rows %u out of range", box_rows);
  GIC_REQUIRE(cols % GEMM_ATOM_K == 0, "tensor-core GEMM operands need K (%llu) to be a multiple of %d", (unsigned long long)cols, GEMM_ATOM_K);
  cuuint64_t gdim[3] = {(cuuint64_t)GEMM_ATOM_K, rows, cols / GEMM_ATOM_K};
  cuuint64_t gstride[2] = {row_stride_elems * 2, (cuuint64_t)GEMM_ATOM_K * 2};
  cuuint32_t box[3] = {(cuuint32_t)GEMM_ATOM_K, box_rows, (cuuint32_t)(GEMM_BLOCK_K / GEMM_ATOM_K)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GIC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return GIC_OK;
}

// Tile width.  Measured on B200 (csrc/microbench.cu): the kernel is bound by how fast one SM can pull operand bytes
// through its TMA ring (~70 GB/s per SM, ~8 TB/s chip-wide from L2), so the best tile minimises
// rounds x bytes-per-tile-per-k-block = ceil(tiles / #SMs) x (A 16 KB + W 128 B x BLOCK_N): wide tiles re-read the
// activation slab less often, narrow ones keep all SMs busy when M is small.
int gemm_bf16_pick_block_n(int M, int N, int split) {
  static const int wide[] = {256, 192, 128, 64, 32};
  static const int narrow[] = {64, 32};  // split (bf16x2) stages carry four operand tiles: two 96 KB stages at 64 wide
  const int* cand = split ? narrow : wide;
  const int n_cand = split ? 2 : 5;
  const long m_tiles = ceil_div(M, GEMM_BLOCK_M);
  const int sms = 148;
  int best = cand[0];
  long best_cost = -1;
  for (int i = 0; i < n_cand; ++i) {
    const int bn = cand[i];
    const long tiles = m_tiles * ceil_div(N, bn);
    const long cost = ((tiles + sms - 1) / sms) * (16384 + 128L * bn);  // relative operand bytes per k-block
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

static int gemm_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int BLOCK_N, bool SPLIT>
static int configure_cfg() {
  GIC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmTile<BLOCK_N, SPLIT>::SMEM_BYTES));
  return GIC_OK;
}

// opt every instantiation into its dynamic shared memory size once (outside any stream capture)
int gemm_bf16_configure() {
  static bool done = false;
  if (done) return GIC_OK;
  GIC_TRY((configure_cfg<32, false>()));
  GIC_TRY((configure_cfg<64, false>()));
  GIC_TRY((configure_cfg<128, false>()));
  GIC_TRY((configure_cfg<192, false>()));
  GIC_TRY((configure_cfg<256, false>()));
  GIC_TRY((configure_cfg<32, true>()));
  GIC_TRY((configure_cfg<64, true>()));
  done = true;
  return GIC_OK;
}

template <int BLOCK_N, bool SPLIT>
static int launch_cfg(const GemmKernelParams& kp, cudaStream_t st) {
  using Tile = GemmTile<BLOCK_N, SPLIT>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, SPLIT>;
  const long tiles = (long)ceil_div(kp.M, GEMM_BLOCK_M) * ceil_div(kp.N, BLOCK_N);
  const int sms = gemm_num_sms();
  dim3 grid((unsigned)(tiles < sms ? tiles : sms));  // persistent: one CTA per SM
  GIC_CHECK_CUDA(launch_kernel(kern, grid, dim3(GEMM_THREADS), (size_t)Tile::SMEM_BYTES, st, kp));
  note_launch();
  return GIC_OK;
}

int launch_gemm_bf16(const GemmBf16Args& a, cudaStream_t st) {
  GIC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm_bf16: empty problem");
  GIC_REQUIRE(a.K % GEMM_ATOM_K == 0, "gemm_bf16: K (%d) must be a multiple of %d", a.K, GEMM_ATOM_K);
  GemmKernelParams kp;
  kp.a_hi = a.a_hi; kp.w_hi = a.w_hi; kp.a_lo = a.a_lo; kp.w_lo = a.w_lo;
  kp.M = a.M; kp.N = a.N; kp.K = a.K; kp.epilogue = a.epilogue; kp.bias = a.bias;
  kp.mma_repeat = a.mma_repeat < 1 ? 1 : a.mma_repeat;
  kp.out_f32 = a.out.f32; kp.ld_f32 = a.ld_out; kp.out_hi = a.out.hi; kp.out_lo = a.out.lo; kp.ld_bf16 = a.ld_out;
  kp.part_val = a.part_val; kp.part_idx = a.part_idx; kp.trace = a.trace;
  GIC_REQUIRE(!(a.epilogue == EPI_RESIDUAL && !a.out.f32), "gemm_bf16: residual epilogue needs the fp32 output");
  if (a.split) {
    switch (a.block_n) {
      case 32: return launch_cfg<32, true>(kp, st);
      case 64: return launch_cfg<64, true>(kp, st);
    }
  } else {
    switch (a.block_n) {
      case 32: return launch_cfg<32, false>(kp, st);
      case 64: return launch_cfg<64, false>(kp, st);
      case 128: return launch_cfg<128, false>(kp, st);
      case 192: return launch_cfg<192, false>(kp, st);
      case 256: return launch_cfg<256, false>(kp, st);
    }
  }
  set_error("gemm_bf16: unsupported block_n %d", a.block_n);
  return GIC_ERR_UNSUPPORTED;
}

}  // namespace gic
