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
// a 4..8-stage TMA->MMA mbarrier ring that keeps streaming across tile boundaries, and two TMEM accumulators so the
// epilogue of tile i overlaps the main loop of tile i+1.  10 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer, warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter on alternate 32-column chunks).
//
// Code-generation rules this file follows (measured round 1, profiles/r1b_gemm_timeline.txt):
//  * the producer and MMA loops run with the WHOLE warp converged and issue through elect.sync.  Under `if (lane == 0)` the
//    compiler cannot prove the TMA / MMA operands warp-uniform and wraps every UTMALDG / UTCHMMA in an
//    ELECT + 5 x R2UR + BRA.U.ANY waterfall: 217 cycles per TMA and 88 per MMA issue (the MMA itself runs 64).
//  * the epilogue kind is a template parameter: a run-time switch inside the per-element loop compiled to real branches
//    (~90 cycles per element with one warp per scheduler; 12 k cycles per 128 x 128 tile against 3 k for its main loop).
//  * epilogue stores go through a per-warp swizzled staging tile so that a warp writes whole 128-byte row segments
//    (lane = row in the TMEM layout would scatter every store instruction over 32 rows).
// BF16X2 ("split") mode: A = A_hi + A_lo, W = W_hi + W_lo (each bf16); three MMAs per k-step
// (hi.hi + hi.lo + lo.hi) into the same accumulator give ~16 mantissa bits.  Round 2: the split mode has the same fused
// layer structure as the bf16 one -- LayerNorm folded in (FOLD), fp16 q/k/v out (OUT_F16), residual + hi/lo copy + row
// statistics (OUT_F32_BF16X2_STATS), wide tiles and CTA pairs -- because it is the mode that meets the north-star caption
// tolerance (>= 99 % of captions identical to the fp32 reference; profiles/r2_precision_screen.jsonl: no 1- or 2-MMA scheme does).
#include <cuda.h>

#include "kernels.cuh"

#ifndef GIC_PAIR_PRELOAD
#define GIC_PAIR_PRELOAD 1  // W preload before griddepcontrol.wait for CTA pairs too (0: single CTAs only)
#endif

namespace gic {


// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (fully active) warp; the compiler keeps the guarded uniform-datapath instructions un-looped
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE 256-row MMA; each stages its own 128 rows of A
// and half of the W tile, so a W byte is delivered to (and read from) shared memory once per pair instead of once per CTA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is counted on the LEADER CTA's mbarrier (same offset; the peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {  // one warp in EACH CTA of the pair, same dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// arrive on the mbarrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
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
  int mma_repeat;  // 1; > 1 only in the microbenchmark probe: re-issue each k-block's MMAs to measure the tensor-pipe rate
  const float* bias;
  float* out_f32; int ld_f32;
  bf16* out_hi; bf16* out_lo; int ld_bf16;
  float* part_val; int* part_idx; int part_ld;  // [M][part_ld]: slot (tile * 2 + column-parity warp) of each row
  float* part_val2;  // EPI_ARGMAX2: the slot's second-best value (candidate generation for the exactly re-scored LM head)
  // EPI_TOPK (retrieval scan, no score matrix): every epilogue thread keeps the GEMM_TOPK_KEEP best (score, column) of the columns it
  // sees for ITS row + the largest score it dropped; written at the end as stream (unit index / m_units) * 2 + column-parity warp of the row
  float* topk_v; int* topk_i; float* topk_u; int topk_streams;
  float* beam_m; float* beam_s;  // EPI_BEAM: per (row, stream) running maximum and sum of exp(score - maximum); KEEP = GEMM_BEAM_KEEP, no topk_u
  // folded LayerNorm on the A operand (see launch_gemm_bf16): out = rstd_r * (acc - mean_r * colsum_n) + bias_n
  const float2* ln_stats; int ln_parts; long ln_stats_ld; int ln_row_mul, ln_row_off; const float* ln_colsum;
  // split-K (see launch_gemm_bf16): `split_k` CTAs share one output tile, each over K / split_k; fp32 partials go to splitk_ws
  // [split_k][M][N] and the CTA that arrives last at splitk_counters[tile] sums them in split order and runs the epilogue
  int split_k; float* splitk_ws; int* splitk_counters;
  int splitk_coop;  // every work item has its own CTA: the K slices of a tile finish it together (see the epilogue)
  int w_static;  // W may be fetched before griddepcontrol.wait (see the kernel)
  StepTrace step_trace;  // in-situ timeline (common.cuh); null buffer = off
  float2* stats_out;  // [ceil(N / 32)][ln_stats_ld] (sum, sum of squares) of the values written, per row and 32-column chunk
  long long* trace;  // null; microbenchmark only: clock64 timeline of CTA 0's producer / MMA / epilogue warps
};

constexpr int GEMM_TOPK_KEEP = 8;
static_assert(GEMM_TOPK_KEEP == GEMM_TOPK_KEEP_PUBLIC, "kernels.cuh publishes the survivor count of the fused top-k epilogue");
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // one 128-byte swizzle atom per row
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_STAGING_BYTES = GEMM_EPI_WARPS * 32 * 128;  // per epilogue warp: 32 rows x 128 B, 16-byte chunks XOR-swizzled

template <int BLOCK_N, bool SPLIT, bool PAIR = false>
struct GemmTile {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // [128 rows][128 B]
  static constexpr int W_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * GEMM_BLOCK_K * 2;  // [BLOCK_N rows][128 B]; a CTA of a pair holds half of them
  static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (SPLIT ? 2 : 1);
  // one persistent CTA per SM: spend (almost) all of its shared memory on the TMA ring
  static constexpr int BUDGET = 227 * 1024 - 1024 /* alignment slack */ - 1024 /* barriers */ - GEMM_STAGING_BYTES;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(STAGES >= 3, "tile does not leave room for a 3-stage ring");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 1024 + GEMM_STAGING_BYTES;
  static constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;   // one accumulator buffer
  static constexpr int TMEM_NEED = 2 * ACC_COLS;                 // double-buffered: epilogue(i) overlaps main loop(i+1)
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;  // power of two
};

template <int EPI>
__device__ __forceinline__ float epi_act(float x, bool precise) {
  if (EPI == EPI_TANH) return tanhf(x);
  if (EPI == EPI_GELU) return precise ? gelu_tanh_sigmoid(x) : gelu_tanh_fast(x);
  if (EPI == EPI_RELU) return fmaxf(x, 0.f);
  return x;
}

// Persistent kernel: grid = min(#tiles, #SMs); CTA c walks tiles c, c + grid, ...  (tile t -> m_tile = t % m_tiles,
// n_tile = t / m_tiles, so the ~148 tiles in flight share few W tiles and each is fetched from HBM once).
// EPI: EPI_NONE / TANH / GELU / RELU (bias + activation), EPI_RESIDUAL (out_f32 += result), EPI_ARGMAX (per-tile row argmax
// partials; optional fp32 logits tap).
// OUT (what the epilogue stores -- compile-time so that the per-element code of one instantiation is a handful of instructions):
enum GemmOut { OUT_NONE = 0 /* argmax partials only */, OUT_F32 = 1, OUT_BF16 = 2, OUT_BF16X2 = 3 /* hi + lo */, OUT_F32_BF16_STATS = 4 /* fused residual:
                fp32 stream + its bf16 copy + per-row (sum, sum of squares) partials for the LayerNorm folded into the next GEMM */,
               OUT_F16 = 5 /* IEEE half through out_hi: q | k | v of the split mode (fp16 KV cache) */,
               OUT_F32_BF16X2_STATS = 6 /* fused residual of the split mode: fp32 stream + hi + lo copies + statistics of hi + lo */ };
// FOLD: LayerNorm folded into this GEMM (see GemmBf16Args::ln_stats).
// RAGGED: N % 32 != 0 or a leading dimension that is not a multiple of 4 -- only these instantiations carry the slow generic
// store path (every kernel here runs once per launch with a cold instruction cache: code size is latency).
// PAIR: launched as clusters of two CTAs that share one 256 x BLOCK_N output tile (cta_group::2, see ptx:: above): CTA rank r owns
// rows [128 r, 128 r + 128) of it (its own A tile, TMEM accumulator and epilogue) and stages W rows [r BLOCK_N / 2, (r + 1) BLOCK_N / 2).
// Only the leader (rank 0) issues MMAs; both CTAs' TMA loads are counted on the leader's full barrier, the leader's commits
// arrive on both CTAs' empty / accumulator-full barriers, and both epilogues release the accumulator on the leader's barrier.
template <int BLOCK_N, bool SPLIT, int EPI, int OUT, bool FOLD, bool RAGGED, bool PAIR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmKernelParams p) {
  using Tile = GemmTile<BLOCK_N, SPLIT, PAIR>;
  static_assert(!PAIR || BLOCK_N % 32 == 0, "CTA pairs: W halves of whole swizzle atoms");
  constexpr int STAGES = Tile::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment; everything below is addressed through 32-bit shared-window
  // addresses derived from this one base so the compiler keeps them in uniform registers
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  constexpr uint32_t OFF_BAR = STAGES * Tile::STAGE_BYTES;
  const uint32_t full_bar = smem_base + OFF_BAR;             // [STAGES]
  const uint32_t empty_bar = full_bar + 8 * STAGES;          // [STAGES]
  const uint32_t tmem_full_bar = empty_bar + 8 * STAGES;     // [2]
  const uint32_t tmem_empty_bar = tmem_full_bar + 16;        // [2]
  const uint32_t tmem_slot = tmem_empty_bar + 16;
  volatile int* s_flag = reinterpret_cast<volatile int*>(smem_gen + OFF_BAR + 512);  // split-K: "this CTA arrived last" broadcast
  float4* s_stage = reinterpret_cast<float4*>(smem_gen + OFF_BAR + 1024);  // [8 warps][32 rows][8 chunks]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  // units along M: single 128-row tiles, or 256-row tiles shared by a CTA pair (this CTA: rows 128 * rank onwards)
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0;
  const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles, m_per_unit = PAIR ? 2 : 1;
  const int total_tiles = m_units * n_tiles;
  const int nk = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  // work item w: tile = w / split_k, K slice = w % split_k -- the slices of a tile are consecutive CTAs (launched, and so resident,
  // together: the cooperative reduction below waits for them)
  const int total_work = total_tiles * p.split_k;
  const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, work_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  pdl_launch_dependents();  // let the next kernel's CTAs be scheduled behind this grid (see common.cuh)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  const long long t_start = tracing ? clock64() : 0;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&p.a_hi);
    ptx::prefetch_tmap(&p.w_hi);
    if (SPLIT) {
      ptx::prefetch_tmap(&p.a_lo);
      ptx::prefetch_tmap(&p.w_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar + 8 * a, 1);
      ptx::mbar_init(tmem_empty_bar + 8 * a, GEMM_EPI_WARPS * (PAIR ? 2 : 1));  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    if (PAIR) ptx::tmem_alloc_pair(tmem_slot, Tile::TMEM_COLS);
    else ptx::tmem_alloc(tmem_slot, Tile::TMEM_COLS);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();  // both CTAs' barriers exist before the peer's TMA / commits / arrivals may touch them
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (tracing && threadIdx.x == 0) p.trace[600] = clock64() - t_start;  // prologue done
  // Weights never depend on the previous kernel: the W halves of the first tile's first ring stages are requested BEFORE
  // griddepcontrol.wait, while the previous kernel's stragglers are still running; after the wait only the A halves are missing
  // (w_static: the caller vouches that W was not written by the kernel launched just before this one)
  uint32_t pre = 0;
  constexpr bool PAIR_PRELOAD = GIC_PAIR_PRELOAD != 0;
  if ((!PAIR || PAIR_PRELOAD) && warp == 0 && p.w_static && work0 < total_work) {
    const int ks0 = work0 % p.split_k;
    const int kb0 = (int)((long)ks0 * nk / p.split_k), nk0 = (int)((long)(ks0 + 1) * nk / p.split_k) - kb0;  // this CTA's first K slice
    pre = (uint32_t)(nk0 < STAGES ? nk0 : STAGES);
    if (ptx::elect_one()) {
      const int n0 = (((work0 / p.split_k) % total_tiles) / m_units) * BLOCK_N;
      for (uint32_t ps = 0; ps < pre; ++ps) {
        const uint32_t fb = full_bar + 8 * ps;
        const uint32_t st = smem_base + ps * Tile::STAGE_BYTES;
        const int kc = (kb0 + (int)ps) * GEMM_BLOCK_K;
        if (PAIR) {  // (both CTAs' barriers exist: the cluster barrier above)
          if (rank == 0) ptx::mbar_expect_tx(fb, 2 * Tile::STAGE_BYTES);
          ptx::tma_load_2d_pair(st + Tile::A_BYTES, &p.w_hi, fb, kc, n0 + (int)rank * (BLOCK_N / 2));
          if (SPLIT) ptx::tma_load_2d_pair(st + 2 * Tile::A_BYTES + Tile::W_BYTES, &p.w_lo, fb, kc, n0 + (int)rank * (BLOCK_N / 2));
        } else {
          ptx::mbar_expect_tx(fb, Tile::STAGE_BYTES);
          ptx::tma_load_2d(st + Tile::A_BYTES, &p.w_hi, fb, kc, n0);
          if (SPLIT) ptx::tma_load_2d(st + 2 * Tile::A_BYTES + Tile::W_BYTES, &p.w_lo, fb, kc, n0);
        }
      }
    }
    __syncwarp();
  }
  pdl_wait();  // prologue above overlapped the previous kernel; its outputs are visible from here on
  const int tslot = trace_begin(p.step_trace, TRACE_GEMM, (BLOCK_N << 4) | (EPI << 1) | (PAIR ? 1 : 0));

  if (warp == 0) {
    // ===== TMA producer: streams k-blocks of successive tiles through the ring without pausing at tile boundaries =====
    uint32_t s = 0, ph = 0, it = 0;
    for (int work = work0; work < total_work; work += work_stride) {
      const int tile = work / p.split_k, ks = work % p.split_k;
      const int m0 = ((tile % m_units) * m_per_unit + (int)rank) * GEMM_BLOCK_M, n0 = (tile / m_units) * BLOCK_N;
      const int kb_end = (int)((long)(ks + 1) * nk / p.split_k);
      for (int kb = (int)((long)ks * nk / p.split_k); kb < kb_end; ++kb, ++it) {
        ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
        if (ptx::elect_one()) {
          const uint32_t st = smem_base + s * Tile::STAGE_BYTES;
          const uint32_t fb = full_bar + 8 * s;
          // stage layout: [A_hi][W_hi] and, in split mode, [A_lo][W_lo] behind them
          if (PAIR && it < pre) {
            ptx::tma_load_2d_pair(st, &p.a_hi, fb, kb * GEMM_BLOCK_K, m0);  // barrier armed and W halves requested before the wait
            if (SPLIT) ptx::tma_load_2d_pair(st + Tile::A_BYTES + Tile::W_BYTES, &p.a_lo, fb, kb * GEMM_BLOCK_K, m0);
          } else if (PAIR) {
            // the leader's full barrier counts the bytes of both CTAs' loads (its arrival is the only pending one, so the
            // phase cannot complete before it has armed the count, even if the peer's bytes land first)
            if (rank == 0) ptx::mbar_expect_tx(fb, 2 * Tile::STAGE_BYTES);
            ptx::tma_load_2d_pair(st, &p.a_hi, fb, kb * GEMM_BLOCK_K, m0);
            ptx::tma_load_2d_pair(st + Tile::A_BYTES, &p.w_hi, fb, kb * GEMM_BLOCK_K, n0 + (int)rank * (BLOCK_N / 2));
            if (SPLIT) {
              ptx::tma_load_2d_pair(st + Tile::A_BYTES + Tile::W_BYTES, &p.a_lo, fb, kb * GEMM_BLOCK_K, m0);
              ptx::tma_load_2d_pair(st + 2 * Tile::A_BYTES + Tile::W_BYTES, &p.w_lo, fb, kb * GEMM_BLOCK_K, n0 + (int)rank * (BLOCK_N / 2));
            }
          } else if (it < pre) {
            ptx::tma_load_2d(st, &p.a_hi, fb, kb * GEMM_BLOCK_K, m0);  // barrier armed and W requested before the wait
            if (SPLIT) ptx::tma_load_2d(st + Tile::A_BYTES + Tile::W_BYTES, &p.a_lo, fb, kb * GEMM_BLOCK_K, m0);
          } else {
            ptx::mbar_expect_tx(fb, Tile::STAGE_BYTES);
            ptx::tma_load_2d(st, &p.a_hi, fb, kb * GEMM_BLOCK_K, m0);
            ptx::tma_load_2d(st + Tile::A_BYTES, &p.w_hi, fb, kb * GEMM_BLOCK_K, n0);
            if (SPLIT) {
              ptx::tma_load_2d(st + Tile::A_BYTES + Tile::W_BYTES, &p.a_lo, fb, kb * GEMM_BLOCK_K, m0);
              ptx::tma_load_2d(st + 2 * Tile::A_BYTES + Tile::W_BYTES, &p.w_lo, fb, kb * GEMM_BLOCK_K, n0);
            }
          }
          if (tracing && it < 60) p.trace[it * 4 + 2] = clock64() - t_start;
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane, whole warp converged), alternating between the two TMEM accumulators =====
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M, BLOCK_N);
    uint32_t s = 0, ph = 0, it = 0, local = 0;
    for (int work = (PAIR && rank != 0) ? total_work : work0; work < total_work; work += work_stride, ++local) {  // (pair: the leader only)
      const int ks = work % p.split_k;
      const int kb_begin = (int)((long)ks * nk / p.split_k), kb_end = (int)((long)(ks + 1) * nk / p.split_k);
      const uint32_t acc = local & 1, use = local >> 1;
      ptx::mbar_wait(tmem_empty_bar + 8 * acc, (use & 1) ^ 1);  // epilogue has drained this accumulator (passes at first use)
      ptx::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * Tile::ACC_COLS;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        ptx::mbar_wait(full_bar + 8 * s, ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          if (tracing && it < 60) p.trace[256 + it * 4 + 1] = clock64() - t_start;
          const uint32_t sa = smem_base + s * Tile::STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc_sw128(sa);
          const uint64_t w_hi = make_smem_desc_sw128(sa + Tile::A_BYTES);
          const uint64_t a_lo = make_smem_desc_sw128(sa + Tile::A_BYTES + Tile::W_BYTES);
          const uint64_t w_lo = make_smem_desc_sw128(sa + 2 * Tile::A_BYTES + Tile::W_BYTES);
          for (int rep = 0; rep < p.mma_repeat; ++rep) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              const uint64_t koff = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per k-step inside the 128-byte swizzle row
              if (PAIR) ptx::umma_bf16_pair(tmem_acc, a_hi + koff, w_hi + koff, idesc, ((kb - kb_begin) | k | rep) != 0);
              else ptx::umma_bf16(tmem_acc, a_hi + koff, w_hi + koff, idesc, ((kb - kb_begin) | k | rep) != 0);
              if (SPLIT && PAIR) {
                ptx::umma_bf16_pair(tmem_acc, a_hi + koff, w_lo + koff, idesc, 1);
                ptx::umma_bf16_pair(tmem_acc, a_lo + koff, w_hi + koff, idesc, 1);
              } else if (SPLIT) {
                ptx::umma_bf16(tmem_acc, a_hi + koff, w_lo + koff, idesc, 1);
                ptx::umma_bf16(tmem_acc, a_lo + koff, w_hi + koff, idesc, 1);
              }
            }
          }
          if (PAIR) {
            ptx::umma_commit_pair(empty_bar + 8 * s);                          // both CTAs' smem slots are free once these MMAs have read them
            if (kb == kb_end - 1) ptx::umma_commit_pair(tmem_full_bar + 8 * acc);  // both CTAs' halves of the accumulator are complete
          } else {
          ptx::umma_commit(empty_bar + 8 * s);                          // smem slot is free once these MMAs have read it
          if (kb == kb_end - 1) ptx::umma_commit(tmem_full_bar + 8 * acc);  // accumulator complete
          }
          if (tracing && it < 60) p.trace[256 + it * 4 + 2] = clock64() - t_start;
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers (lane = row) -> swizzled staging tile -> "coalesced domain" (lane = 16-byte chunk
    // (lane & 7) of rows (lane >> 3) + 4 i) where bias / folded LayerNorm / activation / residual / argmax / row statistics
    // are applied and whole 128-byte row segments are stored.  8 warps: TMEM lane quarter = warp & 3, and the two warps
    // of a quarter take alternate 32-column chunks. =====
    const int e = warp - 2, q = warp & 3, sub = e >> 2;
    float4* stage = s_stage + e * 256;  // this warp's [32 rows][8 x 16 B]
    const bool vec_f32 = (p.ld_f32 % 4 == 0), vec_bf16 = (p.ld_bf16 % 4 == 0);
    const int cchunk = lane & 7, crow0 = lane >> 3;
    constexpr bool fold = FOLD;
    uint32_t local = 0;
    // hand an accumulator back to the MMA warp (pair: the leader's, from both CTAs)
    auto release_acc = [&](uint32_t a) {
      if (PAIR) ptx::mbar_arrive_cluster(tmem_empty_bar + 8 * a, 0);
      else ptx::mbar_arrive(tmem_empty_bar + 8 * a);
    };
    // EPI_TOPK / EPI_BEAM: the launch keeps the row tile of a CTA fixed (grid units a multiple of m_units), so this thread's row never changes
    constexpr bool STREAMS = EPI == EPI_TOPK || EPI == EPI_BEAM;
    constexpr int KEEP = EPI == EPI_BEAM ? GEMM_BEAM_KEEP : GEMM_TOPK_KEEP;
    float tk_v[KEEP], tk_u = -INFINITY;
    int tk_i[KEEP];
    float lse_m = -INFINITY, lse_s = 0.f;  // EPI_BEAM: online log-sum-exp of the columns this thread saw (base 2 inside, natural on output)
    if (STREAMS) {
#pragma unroll
      for (int j = 0; j < KEEP; ++j) { tk_v[j] = -INFINITY; tk_i[j] = -1; }
    }
    for (int work = work0; work < total_work; work += work_stride, ++local) {
      const int tile = work / p.split_k, ks = work % p.split_k;
      const uint32_t acc = local & 1, use = local >> 1;
      const int n_tile = tile / m_units;
      const int m0 = ((tile % m_units) * m_per_unit + (int)rank) * GEMM_BLOCK_M, n0 = n_tile * BLOCK_N;
      const int wrow0 = m0 + q * 32;  // first row of this warp's 32-row band
      if (STREAMS) {
        // ---- retrieval scan / beam-search LM head: scores straight from the TMEM registers (lane = row) into the thread's running top-KEEP; columns arrive
        // in ascending order within a thread, so on equal scores the lower column stays (the exact path's tie rule) ----
        ptx::mbar_wait(tmem_full_bar + 8 * acc, use & 1);
        ptx::tc_fence_after();
        const uint32_t tmem_row = tmem_base + acc * Tile::ACC_COLS + ((uint32_t)(q * 32) << 16);
        if (sub * 32 >= BLOCK_N) {
          ptx::tc_fence_before();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll 1
        for (int c0 = sub * 32; c0 < BLOCK_N; c0 += 64) {
          const int col0 = n0 + c0;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_row + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (c0 + 64 >= BLOCK_N) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          if (col0 >= p.N) continue;  // warp-uniform
          const int ncol = min(32, p.N - col0);
          if (EPI == EPI_BEAM) {
            // online log-sum-exp over the chunk: one rescale per chunk, then 32 independent exp2
            float cm = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) cm = fmaxf(cm, c < ncol ? __uint_as_float(r[c]) : -INFINITY);
            const float nm = fmaxf(lse_m, cm);  // finite: the chunk holds at least one column
            const float nml = nm * 1.4426950408889634f;
            float add = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) add += c < ncol ? exp2f(fmaf(__uint_as_float(r[c]), 1.4426950408889634f, -nml)) : 0.f;
            lse_s = lse_s * exp2f(lse_m * 1.4426950408889634f - nml) + add;  // exp2(-inf) = 0 at the first chunk
            lse_m = nm;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float x = __uint_as_float(r[c]);
            if (c < ncol) {
              if (x > tk_v[KEEP - 1]) {
                tk_u = fmaxf(tk_u, tk_v[KEEP - 1]);
                float pv = x;
                int pi = col0 + c;
#pragma unroll
                for (int j = 0; j < KEEP; ++j)
                  if (pv > tk_v[j]) { const float tv = tk_v[j]; const int ti = tk_i[j]; tk_v[j] = pv; tk_i[j] = pi; pv = tv; pi = ti; }
              } else {
                tk_u = fmaxf(tk_u, x);
              }
            }
          }
        }
        continue;
      }
      if ((EPI == EPI_ARGMAX || EPI == EPI_ARGMAX2) && OUT == OUT_NONE) {
        // ---- LM head on the product path: argmax straight from the TMEM registers (lane = row), nothing is staged or stored.
        // (the staging tile would add ~20 % to the shared-memory traffic that bounds this kernel's main loop) ----
        const int row = wrow0 + lane;
        float mu = 0.f, rs = 1.f;
        if (FOLD) {
          float sx = 0.f, sq = 0.f;
          if (row < p.M) {
            // fixed order (8 interleaved partial sums, then a tree): batch-invariant, and the 8 loads of a round overlap
            const float2* sp = p.ln_stats + (size_t)row * p.ln_row_mul + p.ln_row_off;
            float ax[8], aq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { ax[u] = 0.f; aq[u] = 0.f; }
            for (int part = 0; part < p.ln_parts; part += 8) {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (part + u < p.ln_parts) {
                  const float2 t = __ldcg(sp + (size_t)(part + u) * p.ln_stats_ld);
                  ax[u] += t.x;
                  aq[u] += t.y;
                }
            }
            sx = ((ax[0] + ax[1]) + (ax[2] + ax[3])) + ((ax[4] + ax[5]) + (ax[6] + ax[7]));
            sq = ((aq[0] + aq[1]) + (aq[2] + aq[3])) + ((aq[4] + aq[5]) + (aq[6] + aq[7]));
          }
          const float inv_k = 1.0f / (float)p.K;
          mu = sx * inv_k;
          rs = rsqrtf(fmaxf(sq * inv_k - mu * mu, 0.f) + 1e-5f);
        }
        float bv = -INFINITY, bv2 = -INFINITY;  // best and (EPI_ARGMAX2) second-best value of the slot
        int bi = 0x7fffffff;
        ptx::mbar_wait(tmem_full_bar + 8 * acc, use & 1);
        ptx::tc_fence_after();
        if (tracing && e == 0 && lane == 0 && local < 8) p.trace[512 + local * 4 + 1] = clock64() - t_start;
        const uint32_t tmem_row = tmem_base + acc * Tile::ACC_COLS + ((uint32_t)(q * 32) << 16);
        if (sub * 32 >= BLOCK_N) {
          ptx::tc_fence_before();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll 1
        for (int c0 = sub * 32; c0 < BLOCK_N; c0 += 64) {
          const int col0 = n0 + c0;
          const bool whole = col0 + 32 <= p.N;
          // bias / column sums of the chunk: warp-uniform addresses (one broadcast transaction each), all sixteen requested
          // before the accumulator is read so that their L2 latency is paid once per chunk, not once per use
          float4 b4[8], cs4[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = col0 + 4 * c;
            b4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            cs4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (whole) {
              if (p.bias) b4[c] = __ldg(reinterpret_cast<const float4*>(p.bias + col));
              if (FOLD) cs4[c] = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col));
            } else {
              if (p.bias) {
                if (col < p.N) b4[c].x = __ldg(p.bias + col);
                if (col + 1 < p.N) b4[c].y = __ldg(p.bias + col + 1);
                if (col + 2 < p.N) b4[c].z = __ldg(p.bias + col + 2);
                if (col + 3 < p.N) b4[c].w = __ldg(p.bias + col + 3);
              }
              if (FOLD) {
                if (col < p.N) cs4[c].x = __ldg(p.ln_colsum + col);
                if (col + 1 < p.N) cs4[c].y = __ldg(p.ln_colsum + col + 1);
                if (col + 2 < p.N) cs4[c].z = __ldg(p.ln_colsum + col + 2);
                if (col + 3 < p.N) cs4[c].w = __ldg(p.ln_colsum + col + 3);
              }
            }
          }
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_row + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (c0 + 64 >= BLOCK_N) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          if (col0 >= p.N) continue;  // warp-uniform
          const float nrm = -rs * mu;  // v = rs * acc - rs * mu * colsum + bias
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = col0 + 4 * c;
            float v0 = __uint_as_float(r[4 * c]), v1 = __uint_as_float(r[4 * c + 1]), v2 = __uint_as_float(r[4 * c + 2]), v3 = __uint_as_float(r[4 * c + 3]);
            if (FOLD) {
              v0 = fmaf(rs, v0, fmaf(nrm, cs4[c].x, b4[c].x));
              v1 = fmaf(rs, v1, fmaf(nrm, cs4[c].y, b4[c].y));
              v2 = fmaf(rs, v2, fmaf(nrm, cs4[c].z, b4[c].z));
              v3 = fmaf(rs, v3, fmaf(nrm, cs4[c].w, b4[c].w));
            } else {
              v0 += b4[c].x; v1 += b4[c].y; v2 += b4[c].z; v3 += b4[c].w;
            }
            // columns in increasing order with strict > : the lowest index among equal maxima survives
            bool g;
            if (EPI == EPI_ARGMAX2) {  // also the runner-up value: a displaced best becomes it, otherwise the larger of it and the newcomer
              if (!whole) {
                v0 = col < p.N ? v0 : -INFINITY; v1 = col + 1 < p.N ? v1 : -INFINITY; v2 = col + 2 < p.N ? v2 : -INFINITY; v3 = col + 3 < p.N ? v3 : -INFINITY;
              }
              g = v0 > bv; bv2 = g ? bv : fmaxf(bv2, v0); bv = g ? v0 : bv; bi = g ? col : bi;
              g = v1 > bv; bv2 = g ? bv : fmaxf(bv2, v1); bv = g ? v1 : bv; bi = g ? col + 1 : bi;
              g = v2 > bv; bv2 = g ? bv : fmaxf(bv2, v2); bv = g ? v2 : bv; bi = g ? col + 2 : bi;
              g = v3 > bv; bv2 = g ? bv : fmaxf(bv2, v3); bv = g ? v3 : bv; bi = g ? col + 3 : bi;
            } else {
              g = (whole || col < p.N) && v0 > bv; bv = g ? v0 : bv; bi = g ? col : bi;
              g = (whole || col + 1 < p.N) && v1 > bv; bv = g ? v1 : bv; bi = g ? col + 1 : bi;
              g = (whole || col + 2 < p.N) && v2 > bv; bv = g ? v2 : bv; bi = g ? col + 2 : bi;
              g = (whole || col + 3 < p.N) && v3 > bv; bv = g ? v3 : bv; bi = g ? col + 3 : bi;
            }
          }
        }
        if (row < p.M) {
          const size_t slot = (size_t)row * p.part_ld + (n_tile * 2 + sub);
          p.part_val[slot] = bv;
          p.part_idx[slot] = bi;
          if (EPI == EPI_ARGMAX2) p.part_val2[slot] = bv2;
        }
        if (tracing && e == 0 && lane == 0 && local < 8) p.trace[512 + local * 4 + 2] = clock64() - t_start;
        continue;
      }
      // folded LayerNorm: per-row mean / rstd of the A operand from the producer's per-tile partial sums (sum, sum of squares)
      float mean[8], rstd[8];
      if (fold) {
        // the eight lanes that share a row split the parts between them (part = lane & 7, + 8, ...): parts in the outer loop
        // so that the eight rows' loads of a round are in flight together; fixed order -> the same result in any batch
        float sx[8], sq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sx[i] = 0.f; sq[i] = 0.f; }
        for (int part = cchunk; part < p.ln_parts; part += 8) {
          const float2* sp = p.ln_stats + (size_t)part * p.ln_stats_ld + p.ln_row_off;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = min(wrow0 + crow0 + 4 * i, p.M - 1);
            const float2 t = __ldcg(sp + (size_t)row * p.ln_row_mul);
            sx[i] += t.x;
            sq[i] += t.y;
          }
        }
        const float inv_k = 1.0f / (float)p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            sx[i] += __shfl_xor_sync(0xffffffffu, sx[i], o);
            sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
          }
          const float mu = sx[i] * inv_k;
          mean[i] = mu;
          rstd[i] = rsqrtf(fmaxf(sq[i] * inv_k - mu * mu, 0.f) + 1e-5f);
        }
      }
      float rsum[8], rsq[8], best[8];
      int bidx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { rsum[i] = 0.f; rsq[i] = 0.f; best[i] = -INFINITY; bidx[i] = 0x7fffffff; }
      // operands of a chunk that do not depend on the accumulator (residual in the coalesced layout, bias, column sums) are
      // requested one chunk ahead: the first chunk's before the accumulator wait, chunk c+1's while chunk c is processed
      float4 res_n[8], b4_n = make_float4(0.f, 0.f, 0.f, 0.f), cs4_n = make_float4(0.f, 0.f, 0.f, 0.f);
      auto prefetch = [&](int wrow, int c0) {  // wrow: first row of the 32-row band
        const int col0 = n0 + c0, col = col0 + cchunk * 4;
        b4_n = make_float4(0.f, 0.f, 0.f, 0.f);
        cs4_n = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col0 >= p.N) return;
        if (!RAGGED || ((col0 + 32 <= p.N) && vec_f32 && vec_bf16)) {
          if (EPI == EPI_RESIDUAL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = min(wrow + crow0 + 4 * i, p.M - 1);  // clamped: rows >= M are never stored
              res_n[i] = __ldcg(reinterpret_cast<const float4*>(p.out_f32 + (size_t)row * p.ld_f32 + col));
            }
          }
          if (p.bias) b4_n = __ldg(reinterpret_cast<const float4*>(p.bias + col));  // col is a multiple of 4: 16-byte aligned
          if (FOLD) cs4_n = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col));
        } else {
          if (EPI == EPI_RESIDUAL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = wrow + crow0 + 4 * i;
              res_n[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.M) {
                const float* src = p.out_f32 + (size_t)row * p.ld_f32 + col;
                if (col < p.N) res_n[i].x = __ldcg(src);
                if (col + 1 < p.N) res_n[i].y = __ldcg(src + 1);
                if (col + 2 < p.N) res_n[i].z = __ldcg(src + 2);
                if (col + 3 < p.N) res_n[i].w = __ldcg(src + 3);
              }
            }
          }
          if (p.bias) {
            if (col < p.N) b4_n.x = __ldg(p.bias + col);
            if (col + 1 < p.N) b4_n.y = __ldg(p.bias + col + 1);
            if (col + 2 < p.N) b4_n.z = __ldg(p.bias + col + 2);
            if (col + 3 < p.N) b4_n.w = __ldg(p.bias + col + 3);
          }
          if (FOLD) {
            if (col < p.N) cs4_n.x = __ldg(p.ln_colsum + col);
            if (col + 1 < p.N) cs4_n.y = __ldg(p.ln_colsum + col + 1);
            if (col + 2 < p.N) cs4_n.z = __ldg(p.ln_colsum + col + 2);
            if (col + 3 < p.N) cs4_n.w = __ldg(p.ln_colsum + col + 3);
          }
        }
      };
      // split-K: phase 0 parks this CTA's raw fp32 partial tile in the workspace; phase 1 = the normal epilogue on the sum of all
      // partials (added in split order: deterministic).  Cooperative form (p.splitk_coop: every work item has its own resident CTA):
      // the split_k CTAs of a tile wait for each other at the tile's counter and then EACH finishes its share of the tile's
      // 32 x 32 units (unit u = 4 * column chunk + row quarter belongs to slice u % split_k and to that CTA's warp (u / split_k) % 8),
      // so reduction, epilogue math and stores of a tile are spread over split_k SMs.  Otherwise the CTA that arrives last does all of it.
      const bool split = EPI == EPI_RESIDUAL && p.split_k > 1;  // (the K split is offered for the residual GEMMs only)
      const bool coop = split && p.splitk_coop != 0;
      int* const sk_counter = split ? p.splitk_counters + (PAIR ? tile * 2 + (int)rank : tile) : nullptr;
      const uint32_t tmem_acc = tmem_base + acc * Tile::ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int phase = 0; phase < (split ? 2 : 1); ++phase) {
      const bool to_ws = split && phase == 0, from_ws = split && phase == 1;
      const bool by_unit = from_ws && coop;  // this phase walks the units this CTA owns instead of the warp's TMEM chunks
      if (!to_ws && !by_unit && sub * 32 < BLOCK_N) prefetch(wrow0, sub * 32);
      if (phase == 0) {
        ptx::mbar_wait(tmem_full_bar + 8 * acc, use & 1);
        ptx::tc_fence_after();
        if (tracing && e == 0 && lane == 0 && local < 8) p.trace[512 + local * 4 + 1] = clock64() - t_start;
        if (sub * 32 >= BLOCK_N) {  // BLOCK_N == 32: the second warp of the quarter has no chunk, it only releases the accumulator
          ptx::tc_fence_before();
          if (lane == 0) release_acc(acc);
        }
      }
#pragma unroll 1
      for (int itc = 0;; ++itc) {
        int c0, wrow;  // this iteration's 32-column chunk of the tile and the first row of its 32-row band
        if (by_unit) {
          const int u = ks + p.split_k * (e + GEMM_EPI_WARPS * itc);
          if (u >= 4 * (BLOCK_N / 32)) break;
          c0 = (u >> 2) * 32;
          wrow = m0 + (u & 3) * 32;
          prefetch(wrow, c0);
        } else {
          c0 = sub * 32 + 64 * itc;
          if (c0 >= BLOCK_N) break;
          wrow = wrow0;
        }
        const int col0 = n0 + c0, col = col0 + cchunk * 4;
        const bool tr = tracing && e == 0 && lane == 0 && local == 0 && c0 < 128;
        long long* trc = p.trace + 540 + (c0 >> 6) * 8;
        if (tr) trc[0] = clock64() - t_start;
        // whole chunk inside the matrix and every row pointer 16-byte aligned: the specialised path below
        const bool fast = !RAGGED || ((col0 + 32 <= p.N) && vec_f32 && vec_bf16);
        float4 res[8];
        if (EPI == EPI_RESIDUAL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) res[i] = res_n[i];
        }
        const float4 b4 = b4_n, cs4 = cs4_n;
        if (!to_ws && !by_unit && c0 + 64 < BLOCK_N) prefetch(wrow0, c0 + 64);
        if (!from_ws) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_acc + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (tr) trc[1] = clock64() - t_start;
          if (c0 + 64 >= BLOCK_N) {
            // this warp's last TMEM read of the tile: hand the accumulator back to the MMA warp before the math / stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          if (col0 >= p.N) continue;  // warp-uniform
          // lane = row  ->  staging tile (chunk position XOR row keeps both directions bank-conflict free)
          __syncwarp();  // previous chunk's readers are done with the staging tile
#pragma unroll
          for (int c = 0; c < 8; ++c)
            stage[lane * 8 + (c ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                                                             __uint_as_float(r[4 * c + 3]));
          __syncwarp();
        } else if (col0 >= p.N) continue;
        if (tr) trc[2] = clock64() - t_start;
        if (fast) {
          // ---- specialised path: OUT / EPI / FOLD are compile-time, so this is a few instructions per element ----
          if (tr) trc[3] = clock64() - t_start + (long long)(b4.x * 0.f);
          const size_t rowoff = (size_t)(wrow + crow0);
          constexpr bool W_F32 = OUT == OUT_F32 || OUT == OUT_F32_BF16_STATS || OUT == OUT_F32_BF16X2_STATS;
          constexpr bool W_HI = OUT == OUT_BF16 || OUT == OUT_BF16X2 || OUT == OUT_F32_BF16_STATS || OUT == OUT_F16 || OUT == OUT_F32_BF16X2_STATS;
          constexpr bool W_LO = OUT == OUT_BF16X2 || OUT == OUT_F32_BF16X2_STATS;
          constexpr bool W_STATS = OUT == OUT_F32_BF16_STATS || OUT == OUT_F32_BF16X2_STATS;
          float* pf = W_F32 ? p.out_f32 + rowoff * p.ld_f32 + col : nullptr;
          bf16* ph = W_HI ? p.out_hi + rowoff * p.ld_bf16 + col : nullptr;
          bf16* pl = W_LO ? p.out_lo + rowoff * p.ld_bf16 + col : nullptr;
          const size_t step_f32 = (size_t)4 * p.ld_f32, step_bf = (size_t)4 * p.ld_bf16;
          // three separate passes (load, math, store) so that the eight rows overlap instead of running as eight dependent
          // load -> add -> convert -> store chains
          float4 o[8];
          if (!from_ws) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int lr = crow0 + 4 * i;
              o[i] = stage[lr * 8 + (cchunk ^ (lr & 7))];
            }
          }
          if (split) {
            // (split-K runs on whole, aligned tiles only -- launch_gemm_bf16 checks -- so every chunk takes this path)
            const int rows_ok = min(32, p.M - wrow);
            float* wsp = p.splitk_ws + ((size_t)(wrow + crow0)) * p.N + col;
            const size_t ws_row4 = (size_t)4 * p.N, ws_split = (size_t)p.M * p.N;
            if (to_ws) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (crow0 + 4 * i < rows_ok) __stcg(reinterpret_cast<float4*>(wsp + (size_t)ks * ws_split + i * ws_row4), o[i]);
              continue;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < p.split_k; ++sp) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (crow0 + 4 * i < rows_ok) {
                  const float4 t = __ldcg(reinterpret_cast<const float4*>(wsp + (size_t)sp * ws_split + i * ws_row4));
                  o[i].x += t.x; o[i].y += t.y; o[i].z += t.z; o[i].w += t.w;
                }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (FOLD) {
              o[i].x = rstd[i] * (o[i].x - mean[i] * cs4.x);
              o[i].y = rstd[i] * (o[i].y - mean[i] * cs4.y);
              o[i].z = rstd[i] * (o[i].z - mean[i] * cs4.z);
              o[i].w = rstd[i] * (o[i].w - mean[i] * cs4.w);
            }
            o[i].x = epi_act<EPI>(o[i].x + b4.x, SPLIT);
            o[i].y = epi_act<EPI>(o[i].y + b4.y, SPLIT);
            o[i].z = epi_act<EPI>(o[i].z + b4.z, SPLIT);
            o[i].w = epi_act<EPI>(o[i].w + b4.w, SPLIT);
            if (EPI == EPI_RESIDUAL) { o[i].x += res[i].x; o[i].y += res[i].y; o[i].z += res[i].z; o[i].w += res[i].w; }
          }
          if (tr) trc[5] = clock64() - t_start + (long long)(o[7].w * 0.f);  // math done (residual / bias operands have arrived)
          // rows of this warp's band that exist: all 32 unless this is the ragged last row tile (warp-uniform count)
          const int rows_here = min(32, p.M - wrow);
          if (EPI == EPI_ARGMAX) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // columns in increasing order with strict > : the lowest index among equal maxima survives
              float bv = best[i];
              int bi = bidx[i];
              bool g;
              g = o[i].x > bv; bv = g ? o[i].x : bv; bi = g ? col : bi;
              g = o[i].y > bv; bv = g ? o[i].y : bv; bi = g ? col + 1 : bi;
              g = o[i].z > bv; bv = g ? o[i].z : bv; bi = g ? col + 2 : bi;
              g = o[i].w > bv; bv = g ? o[i].w : bv; bi = g ? col + 3 : bi;
              const bool row_ok = crow0 + 4 * i < rows_here;
              best[i] = row_ok ? bv : best[i];
              bidx[i] = row_ok ? bi : bidx[i];
            }
          }
          uint2 pkh[8], pkl[8];
          if (OUT == OUT_F16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 h01 = __floats2half2_rn(o[i].x, o[i].y), h23 = __floats2half2_rn(o[i].z, o[i].w);
              pkh[i].x = *reinterpret_cast<const uint32_t*>(&h01);
              pkh[i].y = *reinterpret_cast<const uint32_t*>(&h23);
            }
          } else if (W_HI) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(o[i].x, o[i].y), h23 = __floats2bfloat162_rn(o[i].z, o[i].w);
              pkh[i].x = *reinterpret_cast<const uint32_t*>(&h01);
              pkh[i].y = *reinterpret_cast<const uint32_t*>(&h23);
              if (W_LO || W_STATS) {
                float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                if (W_LO) {
                  const __nv_bfloat162 l01 = __floats2bfloat162_rn(o[i].x - f01.x, o[i].y - f01.y), l23 = __floats2bfloat162_rn(o[i].z - f23.x, o[i].w - f23.y);
                  pkl[i].x = *reinterpret_cast<const uint32_t*>(&l01);
                  pkl[i].y = *reinterpret_cast<const uint32_t*>(&l23);
                  if (W_STATS) {  // hi + lo is exact in fp32 (8 + 8 significant bits)
                    const float2 g01 = __bfloat1622float2(l01), g23 = __bfloat1622float2(l23);
                    f01.x += g01.x; f01.y += g01.y; f23.x += g23.x; f23.y += g23.y;
                  }
                }
                if (W_STATS) {
                  // row statistics over the values the next GEMM will actually read: the bf16 (or hi + lo) roundings
                  const bool row_ok = crow0 + 4 * i < rows_here;
                  const float ds = (f01.x + f01.y) + (f23.x + f23.y), dq = (f01.x * f01.x + f01.y * f01.y) + (f23.x * f23.x + f23.y * f23.y);
                  rsum[i] += row_ok ? ds : 0.f;
                  rsq[i] += row_ok ? dq : 0.f;
                }
              }
            }
          }
          if (rows_here == 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (W_F32) *reinterpret_cast<float4*>(pf + i * step_f32) = o[i];
              if (W_HI) *reinterpret_cast<uint2*>(ph + i * step_bf) = pkh[i];
              if (W_LO) *reinterpret_cast<uint2*>(pl + i * step_bf) = pkl[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (crow0 + 4 * i < rows_here) {
                if (W_F32) *reinterpret_cast<float4*>(pf + i * step_f32) = o[i];
                if (W_HI) *reinterpret_cast<uint2*>(ph + i * step_bf) = pkh[i];
                if (W_LO) *reinterpret_cast<uint2*>(pl + i * step_bf) = pkl[i];
              }
            }
          }
          if (tr) trc[6] = clock64() - t_start;  // stores issued
        } else if (RAGGED) {
          // ---- generic path: ragged right edge (col0 + 32 > N) or unaligned leading dimensions ----
          const bool full4 = col + 3 < p.N;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int lr = crow0 + 4 * i, row = wrow + lr;
            float4 o = stage[lr * 8 + (cchunk ^ (lr & 7))];
            if (fold) {
              o.x = rstd[i] * (o.x - mean[i] * cs4.x);
              o.y = rstd[i] * (o.y - mean[i] * cs4.y);
              o.z = rstd[i] * (o.z - mean[i] * cs4.z);
              o.w = rstd[i] * (o.w - mean[i] * cs4.w);
            }
            o.x = epi_act<EPI>(o.x + b4.x, SPLIT);
            o.y = epi_act<EPI>(o.y + b4.y, SPLIT);
            o.z = epi_act<EPI>(o.z + b4.z, SPLIT);
            o.w = epi_act<EPI>(o.w + b4.w, SPLIT);
            if (row >= p.M || col >= p.N) continue;
            if (EPI == EPI_RESIDUAL) { o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w; }
            if (EPI == EPI_ARGMAX) {
              if (o.x > best[i]) { best[i] = o.x; bidx[i] = col; }
              if (col + 1 < p.N && o.y > best[i]) { best[i] = o.y; bidx[i] = col + 1; }
              if (col + 2 < p.N && o.z > best[i]) { best[i] = o.z; bidx[i] = col + 2; }
              if (col + 3 < p.N && o.w > best[i]) { best[i] = o.w; bidx[i] = col + 3; }
            }
            if (p.out_f32) {
              float* dst = p.out_f32 + (size_t)row * p.ld_f32 + col;
              dst[0] = o.x;
              if (col + 1 < p.N) dst[1] = o.y;
              if (col + 2 < p.N) dst[2] = o.z;
              if (col + 3 < p.N) dst[3] = o.w;
            }
            float sv[4] = {o.x, o.y, o.z, o.w};
            if (p.out_hi) {
              bf16* dh = p.out_hi + (size_t)row * p.ld_bf16 + col;
              bf16* dl = p.out_lo ? p.out_lo + (size_t)row * p.ld_bf16 + col : nullptr;
              for (int t = 0; t < 4; ++t)
                if (col + t < p.N) {
                  const bf16 hb = __float2bfloat16_rn(sv[t]);
                  dh[t] = hb;
                  if (dl) dl[t] = __float2bfloat16_rn(sv[t] - __bfloat162float(hb));
                  else sv[t] = __bfloat162float(hb);
                }
            }
            if (OUT == OUT_F32_BF16_STATS) {
              for (int t = 0; t < 4; ++t)
                if (col + t < p.N) { rsum[i] += sv[t]; rsq[i] += sv[t] * sv[t]; }
            }
            (void)full4;
          }
        }
        if (OUT == OUT_F32_BF16_STATS || OUT == OUT_F32_BF16X2_STATS) {
          // one statistics part per 32-column chunk (index = global column / 32): the summation order then depends only on the
          // column, never on the tile width or the batch size, so a row's result is the same in any batch
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              rsum[i] += __shfl_xor_sync(0xffffffffu, rsum[i], o);
              rsq[i] += __shfl_xor_sync(0xffffffffu, rsq[i], o);
            }
            const int row = wrow + crow0 + 4 * i;
            if (cchunk == 0 && row < p.M) p.stats_out[(size_t)(col0 >> 5) * p.ln_stats_ld + row] = make_float2(rsum[i], rsq[i]);
            rsum[i] = 0.f;
            rsq[i] = 0.f;
          }
        }
        if (tr) trc[4] = clock64() - t_start;
      }
      if (to_ws) {
        // publish the partial and count arrivals at this tile
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
        if (coop) {
          // cooperative: wait until every K slice of the tile has been parked (low half of the counter = arrivals); all of them go on
          if (e == 0 && lane == 0) {
            atomicAdd(sk_counter, 1);
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(sk_counter) : "memory");
            } while ((seen & 0xffff) < p.split_k);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          __threadfence();
        } else {
          // only the CTA that arrives last goes on to phase 1
          if (e == 0 && lane == 0) {
            const int old = atomicAdd(sk_counter, 1);
            *s_flag = (old == p.split_k - 1) ? 1 : 0;
            if (old == p.split_k - 1) *sk_counter = 0;  // self-cleaning: ready for the next launch
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int last = *s_flag;
          asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read the flag before a later tile may overwrite it
          if (!last) break;
          __threadfence();
        }
      } else if (by_unit) {
        // departures in the high half; the last CTA to leave has seen every partner finish reading and re-arms the counter for the next launch
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (e == 0 && lane == 0) {
          const int old = atomicAdd(sk_counter, 0x10000);
          if ((old >> 16) == p.split_k - 1) *sk_counter = 0;
        }
      }
      }  // phase
      // per-(tile, column-parity) argmax partials of each row
      if (EPI == EPI_ARGMAX) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[i], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
            if (ov > best[i] || (ov == best[i] && oi < bidx[i])) { best[i] = ov; bidx[i] = oi; }
          }
          const int row = wrow0 + crow0 + 4 * i;
          if (cchunk == 0 && row < p.M) {
            const size_t slot = (size_t)row * p.part_ld + (n_tile * 2 + sub);
            p.part_val[slot] = best[i];
            p.part_idx[slot] = bidx[i];
          }
        }
      }
      if (tracing && e == 0 && lane == 0 && local < 8) p.trace[512 + local * 4 + 2] = clock64() - t_start;
    }
    if (STREAMS && work0 < total_work) {
      const int row = ((work0 % m_units) * m_per_unit + (int)rank) * GEMM_BLOCK_M + q * 32 + lane;
      if (row < p.M && sub * 32 < BLOCK_N) {
        const int stream = (work0 / m_units) * 2 + sub;
        const size_t so = ((size_t)row * p.topk_streams + stream) * KEEP;
#pragma unroll
        for (int j = 0; j < KEEP; ++j) { p.topk_v[so + j] = tk_v[j]; p.topk_i[so + j] = tk_i[j]; }
        if (EPI == EPI_BEAM) {
          p.beam_m[(size_t)row * p.topk_streams + stream] = lse_m;
          p.beam_s[(size_t)row * p.topk_streams + stream] = lse_s;
        } else {
          p.topk_u[(size_t)row * p.topk_streams + stream] = tk_u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  trace_end(p.step_trace, tslot);
  if (PAIR) ptx::cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair's MMAs / remote arrivals may still touch it
  if (warp == 1) {
    if (PAIR) ptx::tmem_dealloc_pair(tmem_base, Tile::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, Tile::TMEM_COLS);
  }
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

// bf16 row-major [rows, cols]: one box = [box_rows][64 cols] lands in shared memory as a 128B-swizzled K-major tile --
// exactly the layout the UMMA descriptors walk.  Out-of-bounds rows / columns are zero-filled.
int make_tma_2d_bf16(TmaDesc* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
  GIC_TRY(tma_init());
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
  GIC_REQUIRE(((uintptr_t)base % 16) == 0 && (row_stride_elems * 2) % 16 == 0,
              "TMA: base must be 16-byte aligned and the row stride a multiple of 8 elements (stride %llu)",
              (unsigned long long)row_stride_elems);
  GIC_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA: box rows %u out of range", box_rows);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)GEMM_BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GIC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return GIC_OK;
}

// Tile width and K split.  Cost model from the round-1 timelines (profiles/): a 64-deep k-block of a 128 x bn tile takes
// 256 + 2 bn cycles (shared-memory port: TMA writes + UMMA operand reads), a CTA's epilogue ~600 + 500 per 64 columns, and a
// split-K tile pays ~3000 for parking / re-reading the partials.  rounds = ceil(work items / #SMs).
// pair (optional): set to 1 when a CTA pair (cta_group::2, 256-row tiles, see the kernel) is expected to be faster; a pair's k-block
// takes max(256 + bn, 2 bn) cycles per CTA (half of the W tile each; the MMA itself needs 2 bn) and there are #SMs / 2 pairs.
// GIC_GEMM_PAIR=0 never pairs, =2 pairs whenever the shape allows it (tests).
void gemm_bf16_pick(int M, int N, int K, int split, int split_k, int* block_n, int* pair) {
  static const int wide[] = {256, 192, 128, 64, 32};
  static const int narrow[] = {128, 64, 32};  // bf16x2 stages carry four operand tiles: 64 KB per stage at 128 columns (3 stages)
  const int* cand = split ? narrow : wide;
  const int n_cand = split ? 3 : 5;
  const long m_tiles = ceil_div(M, GEMM_BLOCK_M);
  const int nk = ceil_div(K, GEMM_BLOCK_K);
  const int sms = cta_limit() > 0 && cta_limit() < 148 ? cta_limit() : 148, S = split_k < 1 ? 1 : split_k;
  // cycles per 64-deep k-block of a 128 x bn tile: the shared-memory port (TMA writes + UMMA operand reads, 256 + 2 bn per operand
  // set) against the tensor pipe (2 bn per MMA; bf16x2: two operand sets, three MMAs)
  auto kb_single = [&](int bn) -> long { return split ? (6L * bn > 2L * (256 + 2 * bn) ? 6L * bn : 2L * (256 + 2 * bn)) : 256 + 2L * bn; };
  long best_cost = -1;
  *block_n = cand[0];
  for (int i = 0; i < n_cand; ++i) {
    const int bn = cand[i];
    const long tiles = m_tiles * ceil_div(N, bn);
    const long rounds = (tiles * S + sms - 1) / sms;
    const long cost = rounds * (ceil_div(nk, S) * kb_single(bn) + 600 + 500L * ceil_div(bn, 64) + (S > 1 ? 3000 : 0));
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; *block_n = bn; }
  }
  if (!pair) return;
  *pair = 0;
  const char* pv = getenv("GIC_GEMM_PAIR");
  const int mode = pv ? atoi(pv) : 1;
  if (mode == 0 || S > 1 || (m_tiles & 1) || cta_limit() > 0) return;
  // measured (profiles/r1ag_microbench.txt): a pair saves ~110 cycles per k-block and costs ~2500 per launch (cluster scheduling + two
  // cluster barriers), so it pays for long K (fc2: 15.4 -> 14.0 us at M = 1024, 59 -> 51 us at M = 10240) and for the many-round LM
  // head (55.9 -> 49.1 us, 1.61 PFLOP/s = this box's cuBLAS burst rate), not for the single-round K = 768 GEMMs (+0.6 us).
  // bf16x2: every shape is a candidate (a pair halves the W bytes of both operand sets); the cost model decides.
  if (mode != 2 && !split && !(nk >= 32 || N >= 8192)) return;
  static const int pair_bn[] = {256, 192, 128, 64};
  long best_pair = -1;
  int bn_pair = 0;
  for (int i = 0; i < 4; ++i) {
    const int bn = pair_bn[i];
    const long units = (m_tiles / 2) * ceil_div(N, bn);
    const long rounds = (units + sms / 2 - 1) / (sms / 2);
    long kb = 256 + bn > 2 * bn ? 256 + bn : 2 * bn;
    if (split) kb = 2L * (256 + bn) > 6L * bn ? 2L * (256 + bn) : 6L * bn;
    const long cost = rounds * ((long)nk * kb + 600 + 500L * ceil_div(bn, 64)) + (split ? 2500 : 0);  // (bf16: the gate above stands in for the launch overhead)
    if (best_pair < 0 || cost < best_pair) { best_pair = cost; bn_pair = bn; }
  }
  if (mode == 2 || best_pair < best_cost) { *pair = 1; *block_n = bn_pair; }
}
// K split of the decode-size residual GEMMs: a function of the SHAPE only (never of M), so that a row's fp32 summation
// order -- and with it its tokens -- does not depend on the batch it is generated in
int gemm_bf16_split_k_for(int N, int K) {
  if (N % 32 != 0) return 1;
  const int nk = ceil_div(K, GEMM_BLOCK_K);
  int S = (nk + 8) / 16;
  return S < 1 ? 1 : (S > 4 ? 4 : S);
}
int gemm_bf16_pick_block_n(int M, int N, int split) {
  int bn;
  gemm_bf16_pick(M, N, 768, split, 1, &bn, nullptr);
  return bn;
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

// The instantiated (EPI, OUT, FOLD) variants; anything else is refused by launch_gemm_bf16.
// X(epilogue, output mode, folded LayerNorm, ragged): fp32 outputs exist in both the aligned and the ragged flavour (the LM
// head's N = 50257, odd test shapes); bf16 outputs are always aligned (every GPT-2 / mapper width is a multiple of 64)
#define GIC_GEMM_VARIANTS_COMMON(X) \
  X(EPI_NONE, OUT_F32, false, false) X(EPI_TANH, OUT_F32, false, false) X(EPI_GELU, OUT_F32, false, false) X(EPI_RELU, OUT_F32, false, false) \
  X(EPI_RESIDUAL, OUT_F32, false, false) X(EPI_ARGMAX, OUT_NONE, false, false) \
  X(EPI_NONE, OUT_F32, false, true) X(EPI_TANH, OUT_F32, false, true) X(EPI_GELU, OUT_F32, false, true) X(EPI_RELU, OUT_F32, false, true) \
  X(EPI_RESIDUAL, OUT_F32, false, true) X(EPI_ARGMAX, OUT_F32, false, true)
#define GIC_GEMM_VARIANTS_BF16(X) \
  X(EPI_NONE, OUT_BF16, false, false) X(EPI_NONE, OUT_BF16, true, false) X(EPI_TANH, OUT_BF16, false, false) X(EPI_GELU, OUT_BF16, false, false) \
  X(EPI_GELU, OUT_BF16, true, false) X(EPI_RELU, OUT_BF16, false, false) X(EPI_RESIDUAL, OUT_F32_BF16_STATS, false, false) \
  X(EPI_ARGMAX, OUT_NONE, true, false) X(EPI_ARGMAX, OUT_F32, true, true) X(EPI_NONE, OUT_F32, true, true) X(EPI_ARGMAX2, OUT_NONE, false, false) \
  X(EPI_BEAM, OUT_NONE, false, false)
#define GIC_GEMM_VARIANTS_SPLIT(X) \
  X(EPI_NONE, OUT_BF16X2, false, false) X(EPI_TANH, OUT_BF16X2, false, false) X(EPI_GELU, OUT_BF16X2, false, false) X(EPI_RELU, OUT_BF16X2, false, false) \
  X(EPI_NONE, OUT_F16, true, false) X(EPI_GELU, OUT_BF16X2, true, false) X(EPI_RESIDUAL, OUT_F32_BF16X2_STATS, false, false)
// the fused GPT-2 layer of the split mode on wide tiles (folded qkv -> fp16, folded fc + GELU -> hi + lo, residual + copies + statistics)
// and the LM head; the narrow tiles (32 / 64) keep the full list above for the mappers and the test hooks
#define GIC_GEMM_VARIANTS_SPLIT_WIDE(X) \
  X(EPI_NONE, OUT_F16, true, false) X(EPI_GELU, OUT_BF16X2, true, false) X(EPI_RESIDUAL, OUT_F32_BF16X2_STATS, false, false) \
  X(EPI_ARGMAX, OUT_NONE, false, false) X(EPI_NONE, OUT_F32, false, false) X(EPI_NONE, OUT_F32, false, true) X(EPI_ARGMAX, OUT_F32, false, true) \
  X(EPI_TOPK, OUT_NONE, false, false) X(EPI_BEAM, OUT_NONE, false, false)

// CTA-pair instantiations (aligned shapes, bf16 operands): the fused decode / prefill GEMMs, the LM head, and the plain fp32-output
// GEMM of the kernel test hook
#define GIC_GEMM_VARIANTS_PAIR(X) \
  X(EPI_NONE, OUT_BF16, true) X(EPI_GELU, OUT_BF16, true) X(EPI_RESIDUAL, OUT_F32_BF16_STATS, false) X(EPI_ARGMAX, OUT_NONE, false) X(EPI_NONE, OUT_F32, false) \
  X(EPI_ARGMAX2, OUT_NONE, false) X(EPI_BEAM, OUT_NONE, false)
#define GIC_GEMM_VARIANTS_PAIR_SPLIT(X) \
  X(EPI_NONE, OUT_F16, true) X(EPI_GELU, OUT_BF16X2, true) X(EPI_RESIDUAL, OUT_F32_BF16X2_STATS, false) X(EPI_ARGMAX, OUT_NONE, false) X(EPI_NONE, OUT_F32, false) \
  X(EPI_TOPK, OUT_NONE, false) X(EPI_BEAM, OUT_NONE, false)

template <int BLOCK_N>
static int configure_pair() {
#define X(E, O, F)                                                                                                                         \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, false, E, O, F, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      GemmTile<BLOCK_N, false, true>::SMEM_BYTES));
  GIC_GEMM_VARIANTS_PAIR(X)
#undef X
#define X(E, O, F)                                                                                                                        \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, true, E, O, F, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      GemmTile<BLOCK_N, true, true>::SMEM_BYTES));
  GIC_GEMM_VARIANTS_PAIR_SPLIT(X)
#undef X
  return GIC_OK;
}

template <int BLOCK_N>
static int configure_split_wide() {
#define X(E, O, F, R)                                                                                                               \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, true, E, O, F, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      GemmTile<BLOCK_N, true>::SMEM_BYTES));
  GIC_GEMM_VARIANTS_SPLIT_WIDE(X)
#undef X
  return GIC_OK;
}

template <int BLOCK_N, bool SPLIT>
static int configure_cfg() {
#define X(E, O, F, R)                                                                                                                \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BLOCK_N, SPLIT, E, O, F, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      GemmTile<BLOCK_N, SPLIT>::SMEM_BYTES));
  GIC_GEMM_VARIANTS_COMMON(X)
  if constexpr (SPLIT) { GIC_GEMM_VARIANTS_SPLIT(X) } else { GIC_GEMM_VARIANTS_BF16(X) }
#undef X
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
  GIC_TRY((configure_split_wide<128>()));
  GIC_TRY((configure_pair<64>()));
  GIC_TRY((configure_pair<128>()));
  GIC_TRY((configure_pair<192>()));
  GIC_TRY((configure_pair<256>()));
  done = true;
  return GIC_OK;
}

// GIC_SPLITK_COOP=0: every K split falls back to the last-arriver reduction (measurement / tests)
static bool splitk_coop_disabled() {
  static const bool off = [] { const char* v = getenv("GIC_SPLITK_COOP"); return v && v[0] == '0'; }();
  return off;
}

static thread_local bool g_gemm_no_pdl = false;  // set by launch_gemm_bf16 from GemmBf16Args::no_pdl for the launch it issues

template <int BLOCK_N, bool SPLIT, int EPI, int OUT, bool FOLD, bool RAGGED>
static int launch_one(const GemmKernelParams& kp, cudaStream_t st) {
  using Tile = GemmTile<BLOCK_N, SPLIT>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, SPLIT, EPI, OUT, FOLD, RAGGED>;
  const long tiles = (long)ceil_div(kp.M, GEMM_BLOCK_M) * ceil_div(kp.N, BLOCK_N) * kp.split_k;

  const int sms = cta_limit() > 0 && cta_limit() < gemm_num_sms() ? cta_limit() : gemm_num_sms();
  dim3 grid((unsigned)(tiles < sms ? tiles : sms));  // persistent: one CTA per SM (of this chain's share, see cta_limit)
  if (EPI == EPI_TOPK || EPI == EPI_BEAM) {  // row-tile-stationary CTAs: the grid is a multiple of the number of row tiles (see launch_gemm_topk_streams)
    const int m_units = ceil_div(kp.M, GEMM_BLOCK_M);
    grid.x = (unsigned)(kp.topk_streams / 2 * m_units);
  }
  GemmKernelParams kq = kp;
  kq.splitk_coop = (kp.split_k > 1 && tiles <= sms && !splitk_coop_disabled()) ? 1 : 0;  // one resident CTA per work item
  if (g_gemm_no_pdl) {
    kern<<<grid, dim3(GEMM_THREADS), (size_t)Tile::SMEM_BYTES, st>>>(kq);
    GIC_CHECK_CUDA(cudaGetLastError());
  } else {
    GIC_CHECK_CUDA(launch_kernel(kern, grid, dim3(GEMM_THREADS), (size_t)Tile::SMEM_BYTES, st, kq));
  }
  note_launch();
  return GIC_OK;
}

// one cluster of two CTAs per 256 x BLOCK_N tile; persistent over as many pairs as can be co-resident (a pair needs both SMs of a TPC)
template <int BLOCK_N, int EPI, int OUT, bool FOLD, bool SPLIT = false>
static int launch_one_pair(const GemmKernelParams& kp, cudaStream_t st) {
  using Tile = GemmTile<BLOCK_N, SPLIT, true>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, SPLIT, EPI, OUT, FOLD, false, true>;
  static int max_pairs = 0;
  if (max_pairs == 0) {
    int n = 0;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(2 * 74); q.blockDim = dim3(GEMM_THREADS); q.dynamicSmemBytes = Tile::SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = gemm_num_sms() / 2; }
    max_pairs = n;
  }
  const long units = (long)(ceil_div(kp.M, GEMM_BLOCK_M) / 2) * ceil_div(kp.N, BLOCK_N) * kp.split_k;
  long pairs = units < max_pairs ? units : max_pairs;
  GemmKernelParams kq = kp;
  kq.splitk_coop = (kp.split_k > 1 && units <= max_pairs && !splitk_coop_disabled()) ? 1 : 0;  // one resident pair per work item
  if (EPI == EPI_TOPK || EPI == EPI_BEAM) pairs = (long)(kp.topk_streams / 2) * (ceil_div(kp.M, GEMM_BLOCK_M) / 2);  // row-tile-stationary pairs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Tile::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled() && !g_gemm_no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  GIC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, kq));
  note_launch();
  return GIC_OK;
}

// (epilogue, output, fold) combinations that exist only in the ragged flavour (test / logits-tap paths)
static bool has_aligned(int epi, int out, bool fold) { return !((epi == EPI_ARGMAX && out == OUT_F32) || (epi == EPI_NONE && out == OUT_F32 && fold)); }

template <int BLOCK_N>
static int launch_cfg_pair(const GemmKernelParams& kp, int epi, int out, bool fold, bool split, cudaStream_t st) {
  if (!split) {
#define X(E, O, F) \
  if (epi == E && out == O && fold == F) return launch_one_pair<BLOCK_N, E, O, F>(kp, st);
    GIC_GEMM_VARIANTS_PAIR(X)
#undef X
  } else {
#define X(E, O, F) \
  if (epi == E && out == O && fold == F) return launch_one_pair<BLOCK_N, E, O, F, true>(kp, st);
    GIC_GEMM_VARIANTS_PAIR_SPLIT(X)
#undef X
  }
  set_error("gemm_bf16: no CTA-pair kernel for epilogue %d with output mode %d%s%s", epi, out, fold ? " + folded LayerNorm" : "", split ? " (bf16x2)" : "");
  return GIC_ERR_UNSUPPORTED;
}

template <int BLOCK_N>
static int launch_cfg_split_wide(const GemmKernelParams& kp, int epi, int out, bool fold, cudaStream_t st) {
  const bool aligned = kp.N % 32 == 0 && kp.ld_f32 % 4 == 0 && kp.ld_bf16 % 4 == 0;
#define X(E, O, F, R) \
  if (epi == E && out == O && fold == F && (R || aligned || O == OUT_NONE) && (!R || !aligned || !has_aligned(E, O, F))) return launch_one<BLOCK_N, true, E, O, F, R>(kp, st);
  GIC_GEMM_VARIANTS_SPLIT_WIDE(X)
#undef X
  set_error("gemm_bf16: no wide bf16x2 kernel for epilogue %d with output mode %d%s", epi, out, fold ? " + folded LayerNorm" : "");
  return GIC_ERR_UNSUPPORTED;
}


template <int BLOCK_N, bool SPLIT>
static int launch_cfg(const GemmKernelParams& kp, int epi, int out, bool fold, cudaStream_t st) {
  // the aligned flavour when the shape allows it and it exists, else the ragged one
  const bool aligned = kp.N % 32 == 0 && kp.ld_f32 % 4 == 0 && kp.ld_bf16 % 4 == 0;
#define X(E, O, F, R) \
  if (epi == E && out == O && fold == F && (R || aligned || O == OUT_NONE) && (!R || !aligned || !has_aligned(E, O, F))) return launch_one<BLOCK_N, SPLIT, E, O, F, R>(kp, st);
  GIC_GEMM_VARIANTS_COMMON(X)
  if constexpr (SPLIT) { GIC_GEMM_VARIANTS_SPLIT(X) } else { GIC_GEMM_VARIANTS_BF16(X) }
#undef X
  set_error("gemm_bf16: no kernel for epilogue %d with output mode %d%s%s", epi, out, fold ? " + folded LayerNorm" : "", SPLIT ? " (bf16x2)" : "");
  return GIC_ERR_UNSUPPORTED;
}

// Streams of the fused top-k epilogue: CTAs (or CTA pairs) are row-tile-stationary -- the grid holds `g` units per row tile, each walking
// every g-th column tile -- and each unit's two column-parity epilogue warps keep one candidate stream per row: 2 g streams per row.
int gemm_topk_streams(int M, int N, int block_n, int pair) {
  const int m_units = pair ? ceil_div(M, GEMM_BLOCK_M) / 2 : ceil_div(M, GEMM_BLOCK_M);
  const int n_tiles = ceil_div(N, block_n);
  int units = pair ? gemm_num_sms() / 2 : gemm_num_sms();
  if (cta_limit() > 0 && cta_limit() < gemm_num_sms()) units = pair ? cta_limit() / 2 : cta_limit();
  int g = units / (m_units < 1 ? 1 : m_units);
  if (g > n_tiles) g = n_tiles;
  if (g < 1) g = 1;
  return 2 * g;
}

int launch_gemm_bf16(const GemmBf16Args& a, cudaStream_t st) {
  struct NoPdlScope { bool prev; explicit NoPdlScope(bool v) : prev(g_gemm_no_pdl) { g_gemm_no_pdl = v; } ~NoPdlScope() { g_gemm_no_pdl = prev; } } no_pdl_scope(a.no_pdl != 0);
  GIC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm_bf16: empty problem");
  GIC_REQUIRE(a.K % 8 == 0, "gemm_bf16: K (%d) must be a multiple of 8", a.K);
  GemmKernelParams kp;
  kp.a_hi = a.a_hi; kp.w_hi = a.w_hi; kp.a_lo = a.a_lo; kp.w_lo = a.w_lo;
  kp.M = a.M; kp.N = a.N; kp.K = a.K; kp.bias = a.bias;
  kp.mma_repeat = a.mma_repeat < 1 ? 1 : a.mma_repeat;
  kp.out_f32 = a.out.f32; kp.ld_f32 = a.ld_out; kp.out_hi = a.out.hi; kp.out_lo = a.out.lo; kp.ld_bf16 = a.ld_out;
  kp.part_val = a.part_val; kp.part_idx = a.part_idx; kp.part_ld = a.part_ld; kp.part_val2 = a.part_val2; kp.trace = a.trace;
  kp.topk_v = a.topk_v; kp.topk_i = a.topk_i; kp.topk_u = a.topk_u; kp.topk_streams = a.topk_streams; kp.beam_m = a.beam_m; kp.beam_s = a.beam_s;
  kp.ln_stats = a.ln_stats; kp.ln_parts = a.ln_parts; kp.ln_stats_ld = a.ln_stats_ld; kp.ln_row_mul = a.ln_row_mul; kp.ln_row_off = a.ln_row_off;
  kp.ln_colsum = a.ln_colsum; kp.stats_out = a.stats_out;
  kp.split_k = a.split_k < 1 ? 1 : a.split_k; kp.splitk_ws = a.splitk_ws; kp.splitk_counters = a.splitk_counters;
  kp.w_static = a.w_static;
  kp.step_trace = trace_desc();
  GIC_REQUIRE(!a.ln_stats || (a.ln_colsum && a.ln_parts > 0), "gemm_bf16: folded LayerNorm needs the column sums and at least one statistics part");
  int epi = a.epilogue;
  if (a.topk_v) {
    GIC_REQUIRE(a.epilogue == EPI_NONE && a.topk_i && (a.topk_u || a.beam_m) && (a.beam_m != nullptr) == (a.beam_s != nullptr) && !a.part_val && !a.out.f32 && !a.out.hi &&
                    !a.ln_stats && kp.split_k == 1,
                "gemm_bf16: the fused top-k epilogue takes raw scores and no other output");
    GIC_REQUIRE(a.topk_streams >= 2 && a.topk_streams == gemm_topk_streams(a.M, a.N, a.block_n, a.pair), "gemm_bf16: top-k stream count %d does not match the launch shape",
                a.topk_streams);
    epi = a.beam_m ? EPI_BEAM : EPI_TOPK;
  }
  if (a.part_val) {
    GIC_REQUIRE(a.epilogue == EPI_NONE && a.part_idx, "gemm_bf16: the fused argmax takes no activation and needs both partial buffers");
    GIC_REQUIRE(a.part_ld >= 2 * ceil_div(a.N, a.block_n), "gemm_bf16: argmax partial rows too short (%d slots for %d)", a.part_ld, 2 * ceil_div(a.N, a.block_n));
    epi = a.part_val2 ? EPI_ARGMAX2 : EPI_ARGMAX;
    GIC_REQUIRE(!a.part_val2 || (!a.out.f32 && !a.out.hi && !a.ln_stats && !a.split), "gemm_bf16: the top-2 argmax epilogue is the plain bf16 head without outputs");
  }
  GIC_REQUIRE(!(epi == EPI_RESIDUAL && !a.out.f32), "gemm_bf16: residual epilogue needs the fp32 output");
  const bool fold = a.ln_stats != nullptr;
  int out = OUT_NONE;
  if (a.stats_out) {
    GIC_REQUIRE(epi == EPI_RESIDUAL && a.out.f32 && a.out.hi, "gemm_bf16: row statistics come with the fused residual epilogue (fp32 + bf16 outputs)");
    GIC_REQUIRE((a.out.lo != nullptr) == (a.split != 0), "gemm_bf16: the fused residual epilogue writes hi + lo copies in bf16x2 mode and a single bf16 copy otherwise");
    GIC_REQUIRE(a.ln_stats_ld >= a.M, "gemm_bf16: statistics leading dimension %ld < M %d", a.ln_stats_ld, a.M);
    out = a.out.lo ? OUT_F32_BF16X2_STATS : OUT_F32_BF16_STATS;
  } else if (a.out_f16) {
    GIC_REQUIRE(a.out.hi && !a.out.lo && !a.out.f32, "gemm_bf16: the fp16 output goes through out.hi alone");
    out = OUT_F16;
  } else if (a.out.f32) {
    GIC_REQUIRE(!a.out.hi && !a.out.lo, "gemm_bf16: fp32 and bf16 outputs together only in the fused residual epilogue");
    out = OUT_F32;
  } else if (a.out.hi) {
    out = a.out.lo ? OUT_BF16X2 : OUT_BF16;
  } else {
    GIC_REQUIRE(epi == EPI_ARGMAX || epi == EPI_ARGMAX2 || epi == EPI_TOPK || epi == EPI_BEAM, "gemm_bf16: no output buffer");
  }
  if (kp.split_k > 1) {
    GIC_REQUIRE(epi != EPI_ARGMAX && epi != EPI_ARGMAX2 && a.splitk_ws && a.splitk_counters, "gemm_bf16: split-K needs its workspace / counters and a storing epilogue");
    GIC_REQUIRE(a.N % 32 == 0 && a.ld_out % 4 == 0, "gemm_bf16: split-K needs N %% 32 == 0 and aligned outputs");
    GIC_REQUIRE(ceil_div(a.K, GEMM_BLOCK_K) >= kp.split_k, "gemm_bf16: more K slices than k-blocks");
  }
  if (a.pair) {
    GIC_REQUIRE(ceil_div(a.M, GEMM_BLOCK_M) % 2 == 0, "gemm_bf16: CTA pairs need an even number of 128-row tiles (M = %d)", a.M);
    GIC_REQUIRE(out == OUT_NONE || (a.N % 32 == 0 && a.ld_out % 4 == 0), "gemm_bf16: CTA pairs need N %% 32 == 0 and aligned outputs");
    switch (a.block_n) {
      case 64: return launch_cfg_pair<64>(kp, epi, out, fold, a.split != 0, st);
      case 128: return launch_cfg_pair<128>(kp, epi, out, fold, a.split != 0, st);
      case 192: return launch_cfg_pair<192>(kp, epi, out, fold, a.split != 0, st);
      case 256: return launch_cfg_pair<256>(kp, epi, out, fold, a.split != 0, st);
    }
    set_error("gemm_bf16: unsupported block_n %d for a CTA pair", a.block_n);
    return GIC_ERR_UNSUPPORTED;
  }
  if (a.split) {
    switch (a.block_n) {
      case 32: return launch_cfg<32, true>(kp, epi, out, fold, st);
      case 64: return launch_cfg<64, true>(kp, epi, out, fold, st);
      case 128: return launch_cfg_split_wide<128>(kp, epi, out, fold, st);
    }
  } else {
    switch (a.block_n) {
      case 32: return launch_cfg<32, false>(kp, epi, out, fold, st);
      case 64: return launch_cfg<64, false>(kp, epi, out, fold, st);
      case 128: return launch_cfg<128, false>(kp, epi, out, fold, st);
      case 192: return launch_cfg<192, false>(kp, epi, out, fold, st);
      case 256: return launch_cfg<256, false>(kp, epi, out, fold, st);
    }
  }
  set_error("gemm_bf16: unsupported block_n %d", a.block_n);
  return GIC_ERR_UNSUPPORTED;
}

}  // namespace gic
