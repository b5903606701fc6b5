// Attention kernels of the caption path (head_dim 64 for GPT-2; 64..160 for the transformer mapper).
//
//  attn_decode  : one new query per (row, head) against the KV cache -- the HBM-bound half of a decode step.
//                 Replaces GPT2Attention.forward's cache `torch.cat` + SDPA for T_q = 1
//                 (HF:models/gpt2/modeling_gpt2.py:185-220, HF:cache_utils.py:102-121).  The new token's K/V are appended to
//                 the cache in the same kernel (no separate cat/copy).  Three kernels:
//                   attn_decode_mma_kernel   bf16 product path: persistent CTAs, per-warp cp.async.bulk + mbarrier ring, S and
//                                            P.V on mma.sync with skewed conflict-free ldmatrix; INDIRECT = beam search through
//                                            an ancestry table instead of a reordered cache
//                   attn_decode_bulk_kernel  the same ring with SIMT math (issue bound; kept for comparison, variants 0 / 1)
//                   attn_decode_kernel       fp32 / bf16x2 modes: one warp per (row, head), 8 / 16 lanes per key with 128-bit
//                                            loads (512 contiguous bytes per warp iteration), fp32 scores, warp-shuffle softmax
//  attn_prefill : causal attention over the P prefix tokens of a row, writing K/V into the cache (prefill);
//                 attn_prefill_mma_kernel (bf16, P <= 64) or attn_seq_kernel.
//  attn_encoder : bidirectional attention of nn.TransformerEncoderLayer's MHA (torch:nn/modules/transformer.py:946-950)
//                 for the transformer mapping network (src/models.py:129-139), head_dim = d/8.
//  kv_reorder   : beam-search cache gather, index_select(0, beam_idx) per layer (HF:cache_utils.py:81-85).
//
// KV cache layout (one layer): K [rows][H][T_max][64], V the same; element type T (bf16 or fp32).
#include <stdlib.h>

#include <atomic>

#include "kernels.cuh"

namespace gic {


constexpr int HD = 64;  // GPT-2 head_dim (small/medium/large: n_embd / n_head = 64)

template <typename T>
__global__ void __launch_bounds__(128) attn_decode_kernel(const T* qkv, T* kcache, T* vcache, ActOut out, const int* d_pos, int rows,
                                                          int H, int t_max, const int* row_map) {
  constexpr int VEC = 16 / sizeof(T);   // elements per 128-bit load
  constexpr int LPK = HD / VEC;         // lanes per key: 8 (bf16) / 16 (fp32)
  constexpr int KPI = 32 / LPK;         // lane groups = keys per warp load instruction
  constexpr int UNR = 4;                // keys per group per batch -> KPI*UNR keys (K and V) in flight per warp, twice (prefetch)
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  pdl_launch_dependents();
  if (wid >= rows * H) return;
  pdl_wait();
  const int row = wid / H, h = wid % H;
  const int d = H * HD;
  // (qkv / d_pos are written by the previous kernels: no __restrict__, so their loads are not invariant and stay below pdl_wait)
  const int pos = __ldcg(d_pos);  // tokens already cached == position of the new token
  const int g = lane / LPK, sub = lane % LPK;
  (void)t_max;
  // compacted batches (finished rows dropped, engine.cu): slot `row` of the activations belongs to cache row row_map[row]; -1 = padding slot
  const int crow = row_map ? __ldcg(row_map + row) : row;
  if (crow < 0) return;

  const T* qrow = qkv + (size_t)row * 3 * d + h * HD;
  T* kbase = kcache + ((size_t)crow * H + h) * t_max * HD;
  T* vbase = vcache + ((size_t)crow * H + h) * t_max * HD;

  // Single pass with an online softmax per lane group (flash-decoding style): K and V of a batch of keys are loaded
  // together and the next batch is requested before the current one is consumed, so a warp keeps 2 x KPI x UNR rows of K
  // and of V in flight; the KPI groups' partial (max, sum, acc) are merged with shuffles at the end.  The new token's
  // K/V come straight from the qkv row (and are appended to the cache by group 0) instead of being read back.
  Vec16<T> qv, knew, vnew;
  qv.load(qrow + sub * VEC);
  knew.load(qrow + d + sub * VEC);
  vnew.load(qrow + 2 * d + sub * VEC);
  Vec16<T> kn[UNR], vn[UNR];
  auto request = [&](int j0) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * KPI + g;
      if (j < pos) {
        kn[u].load(kbase + (size_t)j * HD + sub * VEC);
        vn[u].load(vbase + (size_t)j * HD + sub * VEC);
      }
    }
  };
  request(0);
  if (g == 0) {
    knew.store(kbase + (size_t)pos * HD + sub * VEC);
    vnew.store(vbase + (size_t)pos * HD + sub * VEC);
  }
  float qf[VEC];
  qv.unpack(qf);
#pragma unroll
  for (int i = 0; i < VEC; ++i) qf[i] *= 0.125f;  // 1/sqrt(64), HF :211-220 (sdpa default scale)

  float m = -INFINITY, l = 0.f, acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  auto dot_q = [&](const Vec16<T>& kv) {
    float kf[VEC];
    kv.unpack(kf);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s = fmaf(qf[i], kf[i], s);
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  };
  // the new token (position pos) is handled by group 0
  {
    const float s = dot_q(knew);  // all lanes participate in the shuffles
    if (g == 0) {
      m = s;
      l = 1.f;
      float vf[VEC];
      vnew.unpack(vf);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = vf[i];
    }
  }
  for (int j0 = 0; j0 < pos; j0 += KPI * UNR) {
    Vec16<T> kc[UNR], vc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) { kc[u] = kn[u]; vc[u] = vn[u]; }
    if (j0 + KPI * UNR < pos) request(j0 + KPI * UNR);
    float sc[UNR];
    float bm = m;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const bool ok = (j0 + u * KPI + g) < pos;
      const float s = dot_q(kc[u]);  // shuffles executed by every lane; result ignored where !ok
      sc[u] = ok ? s : -INFINITY;
      bm = fmaxf(bm, sc[u]);
    }
    if (bm > -INFINITY) {
      const float scale = (m == -INFINITY) ? 0.f : expf(m - bm);
      l *= scale;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] *= scale;
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (sc[u] > -INFINITY) {
          const float pr = expf(sc[u] - bm);
          l += pr;
          float vf[VEC];
          vc[u].unpack(vf);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(pr, vf[i], acc[i]);
        }
      }
      m = bm;
    }
  }
  // merge the KPI lane groups
#pragma unroll
  for (int o = LPK; o < 32; o <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const float ol = __shfl_xor_sync(0xffffffffu, l, o);
    const float nm = fmaxf(m, om);
    const float sa = (m == -INFINITY) ? 0.f : expf(m - nm);
    const float sb = (om == -INFINITY) ? 0.f : expf(om - nm);
    l = l * sa + ol * sb;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
      acc[i] = acc[i] * sa + oa * sb;
    }
    m = nm;
  }
  if (g == 0) {
    const float inv = 1.0f / l;
    const size_t o0 = (size_t)row * d + h * HD + sub * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) out.write(o0 + i, acc[i] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 decode attention, bulk-copy pipeline with SIMT math (superseded by attn_decode_mma_kernel below, which keeps this ring).
// One persistent CTA per SM; every warp walks its
// own list of (row, head) items and streams their K / V rows (contiguous pos x 128 bytes each) into its private
// shared-memory ring with cp.async.bulk + mbarrier, up to DEC_STAGES chunks of 32 keys ahead.  The copy engine keeps
// ~100 KB per SM in flight without holding registers, so the HBM stream does not stall on the softmax arithmetic or on
// block scheduling (the one-warp-per-item kernel above pays a ~9 us launch / wave ramp per layer: 23 us for 79 MB).
// ---------------------------------------------------------------------------------------------------------------
// ring geometry: DEC_WARPS warps per CTA, DEC_CHUNK keys per stage, DEC_STAGES stages per warp (~196 KB of smem per SM in each
// of the shapes below; which one runs is picked from the microbenchmark, see launch_attn_decode_bulk)
template <int DEC_WARPS, int DEC_CHUNK, int DEC_STAGES>
struct DecCfg {
  static constexpr int STAGE_BYTES = 2 * DEC_CHUNK * HD * 2;  // K chunk + V chunk, bf16
  static constexpr int SMEM_BYTES = DEC_WARPS * DEC_STAGES * STAGE_BYTES + DEC_WARPS * DEC_STAGES * 8 + 128;
};

__device__ __forceinline__ uint32_t dec_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dec_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dec_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

template <int DEC_WARPS, int DEC_CHUNK, int DEC_STAGES>
__global__ void __launch_bounds__(DEC_WARPS * 32, 1) attn_decode_bulk_kernel(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out,
                                                                           const int* d_pos, int rows, int H, int t_max) {
  constexpr int DEC_STAGE_BYTES = DecCfg<DEC_WARPS, DEC_CHUNK, DEC_STAGES>::STAGE_BYTES;
  extern __shared__ uint8_t dec_smem_raw[];
  const uint32_t smem_base = (dec_smem_u32(dec_smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_base + warp * (DEC_STAGES * DEC_STAGE_BYTES);
  const uint32_t bars = smem_base + DEC_WARPS * DEC_STAGES * DEC_STAGE_BYTES + warp * (DEC_STAGES * 8);
  pdl_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < DEC_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_wait();
  const int pos = __ldcg(d_pos);  // tokens already cached == position of the new token
  const int d = H * HD;
  const int n_items = rows * H;
  const int wstride = gridDim.x * DEC_WARPS;
  const int w0 = blockIdx.x * DEC_WARPS + warp;
  const int nch = (pos + DEC_CHUNK - 1) / DEC_CHUNK;  // chunks per item (0 when nothing is cached yet)
  const int my_items = w0 < n_items ? (n_items - 1 - w0) / wstride + 1 : 0;
  const int total_chunks = my_items * nch;
  const int g = lane >> 3, sub = lane & 7;  // 4 lane groups of 8: a group takes keys g, g + 4, ... of a chunk

  // producer side (lane 0): chunk n of this warp's stream -> stage n % DEC_STAGES
  auto issue = [&](int n) {
    const int item = w0 + (n / nch) * wstride, c = n % nch;
    const int row = item / H, h = item % H;
    const int nkeys = min(DEC_CHUNK, pos - c * DEC_CHUNK);
    const uint32_t bytes = (uint32_t)nkeys * HD * 2;
    const int s = n % DEC_STAGES;
    const size_t off = (((size_t)row * H + h) * t_max + (size_t)c * DEC_CHUNK) * HD;
    const uint32_t bar = bars + 8 * s, dst = ring + s * DEC_STAGE_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * bytes) : "memory");
    dec_bulk_load(dst, kcache + off, bytes, bar);
    dec_bulk_load(dst + DEC_CHUNK * HD * 2, vcache + off, bytes, bar);
  };
  if (lane == 0)
    for (int n = 0; n < DEC_STAGES - 1 && n < total_chunks; ++n) issue(n);

  int n = 0;
  // q / k / v of the new token are requested one item ahead (they come from L2: the qkv GEMM has just written them)
  Vec16<bf16> qv_n, knew_n, vnew_n;
  auto load_new = [&](int item) {
    const bf16* qrow = qkv + (size_t)(item / H) * 3 * d + (item % H) * HD;
    qv_n.load(qrow + sub * 8);
    knew_n.load(qrow + d + sub * 8);
    vnew_n.load(qrow + 2 * d + sub * 8);
  };
  if (my_items > 0) load_new(w0);
  for (int ii = 0; ii < my_items; ++ii) {
    const int item = w0 + ii * wstride;
    const int row = item / H, h = item % H;
    const Vec16<bf16> qv = qv_n, knew = knew_n, vnew = vnew_n;
    if (ii + 1 < my_items) load_new(item + wstride);
    float qf[8];
    qv.unpack(qf);
#pragma unroll
    for (int i = 0; i < 8; ++i) qf[i] *= 0.125f;  // 1/sqrt(64), HF :211-220 (sdpa default scale)
    // the new token (position pos): its K / V come straight from the qkv row and are appended to the cache by group 0
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    {
      float kf[8];
      knew.unpack(kf);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(qf[i], kf[i], s);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (g == 0) {
        m = s;
        l = 1.f;
        vnew.unpack(acc);
        bf16* kdst = kcache + (((size_t)row * H + h) * t_max + pos) * HD + sub * 8;
        bf16* vdst = vcache + (((size_t)row * H + h) * t_max + pos) * HD + sub * 8;
        knew.store(kdst);
        vnew.store(vdst);
      }
    }
    for (int c = 0; c < nch; ++c, ++n) {
      // keep the ring full: the stage freed by the previous chunk takes chunk n + DEC_STAGES - 1
      if (lane == 0 && n + DEC_STAGES - 1 < total_chunks) issue(n + DEC_STAGES - 1);
      const int s = n % DEC_STAGES;
      dec_mbar_wait(bars + 8 * s, (uint32_t)((n / DEC_STAGES) & 1));
      const int nkeys = min(DEC_CHUNK, pos - c * DEC_CHUNK);
      const uint32_t kst = ring + s * DEC_STAGE_BYTES + sub * 16, vst = kst + DEC_CHUNK * HD * 2;
      float sc[DEC_CHUNK / 4];
      float bm = m;
#pragma unroll
      for (int u = 0; u < DEC_CHUNK / 4; ++u) {
        const int j = u * 4 + g;
        Vec16<bf16> kv;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(kv.raw.x), "=r"(kv.raw.y), "=r"(kv.raw.z), "=r"(kv.raw.w) : "r"(kst + j * (HD * 2)));
        float kf[8];
        kv.unpack(kf);
        float sdot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sdot = fmaf(qf[i], kf[i], sdot);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
        sc[u] = j < nkeys ? sdot : -INFINITY;  // rows beyond nkeys hold stale bytes of an earlier chunk: masked here
        bm = fmaxf(bm, sc[u]);
      }
      if (bm > -INFINITY) {
        const float scale = (m == -INFINITY) ? 0.f : expf(m - bm);
        l *= scale;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= scale;
#pragma unroll
        for (int u = 0; u < DEC_CHUNK / 4; ++u) {
          const int j = u * 4 + g;
          if (j < nkeys) {
            const float pr = expf(sc[u] - bm);
            l += pr;
            Vec16<bf16> vv;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(vv.raw.x), "=r"(vv.raw.y), "=r"(vv.raw.z), "=r"(vv.raw.w) : "r"(vst + j * (HD * 2)));
            float vf[8];
            vv.unpack(vf);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, vf[i], acc[i]);
          }
        }
        m = bm;
      }
      __syncwarp();  // every lane is done reading this stage before lane 0 may refill it (next iteration's issue)
    }
    // merge the 4 lane groups
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const float ol = __shfl_xor_sync(0xffffffffu, l, o);
      const float nm = fmaxf(m, om);
      const float sa = (m == -INFINITY) ? 0.f : expf(m - nm);
      const float sb = (om == -INFINITY) ? 0.f : expf(om - nm);
      l = l * sa + ol * sb;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
        acc[i] = acc[i] * sa + oa * sb;
      }
      m = nm;
    }
    if (g == 0) {
      const float inv = 1.0f / l;
      float of[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) of[i] = acc[i] * inv;
      Vec16<bf16> ov;
      ov.pack(of);
      ov.store(out + (size_t)row * d + h * HD + sub * 8);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// bf16 decode attention on the warp-level tensor cores (mma.sync m16n8k16), same bulk-copy ring as above.
// The SIMT kernel above executes ~530 warp instructions per 16-key chunk (unpack + FMA + shuffle trees) and is ISSUE bound:
// ncu shows 67 % issue-slot utilisation with 4 warps per scheduler at 4.3 TB/s, against 5.9 TB/s for a plain streaming read of
// the same bytes (profiles/r1z_attn_*.txt).  Here a chunk is 4 ldmatrix + 8 mma for S = q.K^T and 4 ldmatrix.trans + 16 mma
// for O += P.V (P split into bf16 hi + lo so that the probabilities keep ~16 mantissa bits).
//
// The bulk copy lays a chunk out linearly ([key][64 dims], 128-byte rows), which would make every ldmatrix a 16-way bank
// conflict (8 rows 128 bytes apart).  Instead of a swizzled copy, the MATRIX ROWS are skewed: the row of key r reads 16-byte
// column (j ^ (r & 7)), so the eight rows of a matrix hit all 32 banks.  The skew is undone by the other operand:
//   S:  A row g holds q with its 16-byte columns permuted by ^g, so C[g][n] is a true dot product only for n == g: the
//       score of key g (n-tile 0) and key 8 + g (n-tile 1) sit on the diagonal, in lane (g, t = g >> 1), register g & 1.
//   O:  A = P with P[k] placed in row (k & 7) only -- exactly the diagonal positions the scores came out in, so P never moves
//       between lanes.  Accumulator j of row r then holds the partial output of dims column (j ^ r) from keys = r mod 8, and
//       a 3-step exchange at the end of the item (14 shuffles) sums the 8 rows while un-skewing: lane (g, t) ends with dims
//       8 g + 2 t, + 1, i.e. the warp stores one coalesced 128-byte row.
// The new token's K / V row is written into the free slot behind the cached keys of the item's last chunk (and appended to
// the cache), so all pos + 1 keys go through the same path.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D (+)= A . B, A rows 8..15 are zero (a1 = a3 = 0): only c0, c1 (row = lane / 4) are meaningful.
// F16: the operands are IEEE half (q / K / V / P of the bf16x2 engine, whose KV cache is fp16: 11 significant bits in the same
// 2 bytes -- rounding q / k / v to bf16 alone costs 6 % of the captions on random-init weights, to fp16 none,
// profiles/r2_precision_screen.jsonl KVonly_*)
template <bool F16>
__device__ __forceinline__ void mma_16816_top(float& c0, float& c1, float& c2, float& c3, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                 : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                 : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-bit operand pair in the kernel's element type, and the value a single element rounds to
template <bool F16> __device__ __forceinline__ uint32_t pack_op2(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
template <bool F16> __device__ __forceinline__ float round_op(float v) { return F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v)); }
// q * 1/8 on a packed pair (1/sqrt(64), HF :211-220 sdpa default scale; exact in both formats short of underflow)
template <bool F16> __device__ __forceinline__ uint32_t scale_eighth(uint32_t q) {
  if (F16) {
    const __half2 v = __hmul2(*reinterpret_cast<const __half2*>(&q), __floats2half2_rn(0.125f, 0.125f));
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __nv_bfloat162 v = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&q), __floats2bfloat162_rn(0.125f, 0.125f));
  return *reinterpret_cast<const uint32_t*>(&v);
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr int MMA_CHUNK = 16;                          // keys per stage
constexpr int MMA_STAGE_BYTES = 2 * MMA_CHUNK * HD * 2;  // K chunk + V chunk
template <int WARPS, int STAGES>
struct DecMmaCfg { static constexpr int SMEM_BYTES = WARPS * STAGES * MMA_STAGE_BYTES + WARPS * STAGES * 8 + 128; };

// INDIRECT (beam search): the cache is never reordered.  Row r's first n_prefix positions are read from the prefill row of its image
// ((r / beams) * beams), position n_prefix + g from cache row anc[r * anc_ld + g] -- the row that generated the g-th token of r's
// current hypothesis (beam_ancestry_kernel) -- and the new token is appended to r's own row.  This replaces HF's
// DynamicCache.reorder_cache (HF:cache_utils.py:81-85), a 2 x 28 GB gather per step at config 3.
// F16: q / k / v and the cache are IEEE half (typed bf16* here: 2-byte elements either way) and the output is written as a bf16
// hi + lo pair (out, out_lo) -- the A operand of the bf16x2 c_proj GEMM.
template <int WARPS, int STAGES, bool INDIRECT, bool F16 = false>
__global__ void __launch_bounds__(WARPS * 32, 1) attn_decode_mma_kernel(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, const int* d_pos,
                                                                      int rows, int H, int t_max, const int* anc, int anc_ld, int n_prefix, int beams, StepTrace step_trace,
                                                                      bf16* out_lo, const int* row_map) {
  extern __shared__ uint8_t dec_smem_raw[];
  const uint32_t smem_base = (dec_smem_u32(dec_smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_base + warp * (STAGES * MMA_STAGE_BYTES);
  const uint32_t bars = smem_base + WARPS * STAGES * MMA_STAGE_BYTES + warp * (STAGES * 8);
  pdl_launch_dependents();
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero this warp's ring: rows of a stage that no copy has filled yet are multiplied by P = 0 and must not hold NaN / Inf
  for (int i = lane; i < STAGES * MMA_STAGE_BYTES / 16; i += 32)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ring + i * 16), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above before the bulk copies into the same bytes
  __syncwarp();
  pdl_wait();
  const int tslot = trace_begin(step_trace, TRACE_ATTN_DECODE, 0);
  const int pos = __ldcg(d_pos);  // tokens already cached == position of the new token
  const int d = H * HD;
  const int n_items = rows * H;
  const int wstride = gridDim.x * WARPS;
  const int w0 = blockIdx.x * WARPS + warp;
  const int nch = pos / MMA_CHUNK + 1;  // chunks covering the pos cached keys + the new one
  const int my_items = w0 < n_items ? (n_items - 1 - w0) / wstride + 1 : 0;
  const int total_chunks = my_items * nch;
  const int g = lane >> 2, t = lane & 3;

  // compacted batches (finished rows dropped, engine.cu): activation slot r belongs to cache row row_map[r]; -1 = padding slot, which
  // walks cache row 0 like any other item (the ring's bookkeeping stays uniform) but appends nothing
  auto cache_item = [&](int item) -> int {
    if (INDIRECT || row_map == nullptr) return item;
    const int r = item / H, c = __ldcg(row_map + r);
    return (c < 0 ? 0 : c) * H + (item - r * H);
  };
  // producer side (lane 0): chunk n of this warp's stream -> stage n % STAGES (the cached keys of the chunk only).
  // K / V of item i start at i * t_max * 64 elements of their plane (item = row * H + head).
  int p_item = w0, p_c = 0, p_s = 0;  // producer cursor: item, chunk of the item, stage
  // direct: lane 0 issues; INDIRECT: the whole warp calls it (lane 0: barrier + the contiguous prefix part, lane l: generated key l)
  auto issue_next = [&]() {
    const int nkeys = min(MMA_CHUNK, pos - p_c * MMA_CHUNK);  // cached keys in this chunk (0..16)
    const uint32_t bytes = (uint32_t)nkeys * HD * 2;
    const uint32_t bar = bars + 8 * p_s, dst = ring + p_s * MMA_STAGE_BYTES;
    if (!INDIRECT) {
      const size_t off = ((size_t)cache_item(p_item) * t_max + (size_t)p_c * MMA_CHUNK) * HD;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * bytes) : "memory");
      if (bytes > 0) {
        dec_bulk_load(dst, kcache + off, bytes, bar);
        dec_bulk_load(dst + MMA_CHUNK * HD * 2, vcache + off, bytes, bar);
      }
    } else {
      const int prow = p_item / H, ph = p_item - prow * H;
      const int k0 = p_c * MMA_CHUNK;                        // first key of the chunk
      const int npre = max(0, min(nkeys, n_prefix - k0));    // ... of which this many come from the image's prefill row
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * bytes) : "memory");
        if (npre > 0) {
          const size_t off = ((((size_t)(prow / beams) * beams) * H + ph) * t_max + k0) * HD;
          dec_bulk_load(dst, kcache + off, (uint32_t)npre * HD * 2, bar);
          dec_bulk_load(dst + MMA_CHUNK * HD * 2, vcache + off, (uint32_t)npre * HD * 2, bar);
        }
      }
      const int j = npre + lane;  // slot of the chunk served by this lane
      if (j < nkeys) {
        const int src_row = __ldcg(anc + (size_t)prow * anc_ld + (k0 + j - n_prefix));
        const size_t off = ((((size_t)src_row) * H + ph) * t_max + k0 + j) * HD;
        dec_bulk_load(dst + j * (HD * 2), kcache + off, HD * 2, bar);
        dec_bulk_load(dst + MMA_CHUNK * HD * 2 + j * (HD * 2), vcache + off, HD * 2, bar);
      }
    }
    if (++p_c == nch) { p_c = 0; p_item += wstride; }
    if (++p_s == STAGES) p_s = 0;
  };
  int issued = 0;
  if (INDIRECT || lane == 0)
    for (; issued < STAGES - 1 && issued < total_chunks; ++issued) issue_next();

  // q in A-fragment order for row g (16-byte column j of q sits where column j ^ g is expected) and the new token's k / v
  // (16 bytes per lane & 7), requested one item ahead
  uint32_t qa_n[8];
  uint4 knew_n, vnew_n;
  auto load_new = [&](int item) {
    const int row = item / H;
    const bf16* qrow = qkv + (size_t)item * HD + (size_t)row * 2 * d;  // row * 3 d + head * 64
    const uint32_t* q32 = reinterpret_cast<const uint32_t*>(qrow);
#pragma unroll
    for (int j = 0; j < 8; ++j) qa_n[j] = __ldcg(q32 + 4 * (j ^ g) + t);
    knew_n = __ldcg(reinterpret_cast<const uint4*>(qrow + d) + (lane & 7));
    vnew_n = __ldcg(reinterpret_cast<const uint4*>(qrow + 2 * d) + (lane & 7));
  };
  if (my_items > 0) load_new(w0);
  // ldmatrix row addresses inside a stage (r = lane & 7 is the matrix row, m = lane >> 3 the matrix of the x4):
  //   K (plain): key 8 (m >> 1) + r, 16-byte column (2 ks + (m & 1)) ^ r   ->  (b0, b1) of n-tile 0, then of n-tile 1
  //   V (trans): key 8 (m & 1) + r, 16-byte column (2 jj + (m >> 1)) ^ r   ->  (b0, b1) of accumulator 2 jj, then of 2 jj + 1
  const int r8 = lane & 7, mlo = (lane >> 3) & 1, mhi = lane >> 4;
  const uint32_t k_row_off = (uint32_t)((8 * mhi + r8) * (HD * 2));
  const uint32_t v_row_off = (uint32_t)(MMA_CHUNK * HD * 2 + (8 * mlo + r8) * (HD * 2));
  const bool diag = t == (g >> 1);  // this lane holds the scores of keys g and 8 + g of a chunk
  const bool odd = g & 1;
  int c_s = 0;
  uint32_t c_ph = 0;
  for (int ii = 0; ii < my_items; ++ii) {
    const int item = w0 + ii * wstride;
    const bool live_slot = INDIRECT || row_map == nullptr || __ldcg(row_map + item / H) >= 0;
    const int citem = cache_item(item);
    uint32_t qa[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qa[i] = scale_eighth<F16>(qa_n[i]);
    const uint4 knew = knew_n, vnew = vnew_n;
    if (ii + 1 < my_items) load_new(item + wstride);
    float m = -INFINITY, l = 0.f;  // running maximum (warp-uniform) and this lane's share of the running sum
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }
    for (int c = 0; c < nch; ++c) {
      // keep the ring full: the stage freed by the previous chunk takes the next chunk of the stream
      if ((INDIRECT || lane == 0) && issued < total_chunks) { issue_next(); ++issued; }
      const uint32_t st = ring + c_s * MMA_STAGE_BYTES;
      const int cached = min(MMA_CHUNK, pos - c * MMA_CHUNK);  // cached keys of this chunk
      const bool last = c == nch - 1;                            // ... followed by the new token in slot `cached` (< 16 here)
      if (last && lane < 8) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + cached * (HD * 2) + lane * 16), "r"(knew.x), "r"(knew.y), "r"(knew.z), "r"(knew.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + MMA_CHUNK * HD * 2 + cached * (HD * 2) + lane * 16), "r"(vnew.x), "r"(vnew.y), "r"(vnew.z), "r"(vnew.w) : "memory");
        const size_t coff = ((size_t)citem * t_max + pos) * HD + lane * 8;
        if (live_slot) {
          *reinterpret_cast<uint4*>(kcache + coff) = knew;  // append to the cache (HF:cache_utils.py:102-121)
          *reinterpret_cast<uint4*>(vcache + coff) = vnew;
        }
      }
      dec_mbar_wait(bars + 8 * c_s, c_ph);
      __syncwarp();  // the new token's row is visible to every lane's ldmatrix
      const int nkeys = cached + (last ? 1 : 0);
      // ---- S = q . K^T over the 16 key slots (diagonal entries only) ----
      // (two accumulator sets per n-tile: dependent chains of two MMAs instead of four)
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, u0[4] = {0.f, 0.f, 0.f, 0.f}, u1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ks += 2) {
        uint32_t b00, b01, b10, b11, c00, c01, c10, c11;
        ldsm_x4(st + k_row_off + 16 * ((2 * ks + mlo) ^ r8), b00, b01, b10, b11);
        ldsm_x4(st + k_row_off + 16 * ((2 * ks + 2 + mlo) ^ r8), c00, c01, c10, c11);
        mma_16816_top<F16>(s0[0], s0[1], s0[2], s0[3], qa[2 * ks], qa[2 * ks + 1], b00, b01);
        mma_16816_top<F16>(s1[0], s1[1], s1[2], s1[3], qa[2 * ks], qa[2 * ks + 1], b10, b11);
        mma_16816_top<F16>(u0[0], u0[1], u0[2], u0[3], qa[2 * ks + 2], qa[2 * ks + 3], c00, c01);
        mma_16816_top<F16>(u1[0], u1[1], u1[2], u1[3], qa[2 * ks + 2], qa[2 * ks + 3], c10, c11);
      }
      // slots >= nkeys hold stale bytes -> masked by selection
      float sv0 = odd ? s0[1] + u0[1] : s0[0] + u0[0], sv1 = odd ? s1[1] + u1[1] : s1[0] + u1[0];  // keys g and 8 + g (meaningful on the diagonal lanes)
      if (!diag || g >= nkeys) sv0 = -INFINITY;
      if (!diag || 8 + g >= nkeys) sv1 = -INFINITY;
      float cm = fmaxf(sv0, sv1);
      asm volatile("redux.sync.max.f32 %0, %0, 0xffffffff;" : "+f"(cm));  // sm_100a: one warp-wide float reduction instead of five shuffles
      const float mn = fmaxf(m, cm);                 // finite: every chunk holds at least one valid key
      const float mnl = mn * LOG2E;
      const float scale = exp2f(m * LOG2E - mnl);    // 0 at the first chunk (m = -inf)
      m = mn;
      const float p0 = exp2f(fmaf(sv0, LOG2E, -mnl)), p1 = exp2f(fmaf(sv1, LOG2E, -mnl));  // 0 off the diagonal and for masked slots
      l = l * scale + (p0 + p1);
      // P as bf16 hi + lo, in row g at k = g (a0) and k = 8 + g (a2): element g & 1 of the register pair
      const float p0h = round_op<F16>(p0), p1h = round_op<F16>(p1);
      const uint32_t a0h = odd ? pack_op2<F16>(0.f, p0h) : pack_op2<F16>(p0h, 0.f), a2h = odd ? pack_op2<F16>(0.f, p1h) : pack_op2<F16>(p1h, 0.f);
      const uint32_t a0l = odd ? pack_op2<F16>(0.f, p0 - p0h) : pack_op2<F16>(p0 - p0h, 0.f), a2l = odd ? pack_op2<F16>(0.f, p1 - p1h) : pack_op2<F16>(p1 - p1h, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j][0] *= scale; o[j][1] *= scale; }
      // ---- O += P . V (accumulator j of row r: dims column j ^ r, keys = r mod 8) ----
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t v00, v01, v10, v11;
        ldsm_x4_trans(st + v_row_off + 16 * ((2 * jj + mhi) ^ r8), v00, v01, v10, v11);
        mma_16816_top<F16>(o[2 * jj][0], o[2 * jj][1], o[2 * jj][2], o[2 * jj][3], a0h, a2h, v00, v01);
        mma_16816_top<F16>(o[2 * jj + 1][0], o[2 * jj + 1][1], o[2 * jj + 1][2], o[2 * jj + 1][3], a0h, a2h, v10, v11);
        mma_16816_top<F16>(o[2 * jj][0], o[2 * jj][1], o[2 * jj][2], o[2 * jj][3], a0l, a2l, v00, v01);
        mma_16816_top<F16>(o[2 * jj + 1][0], o[2 * jj + 1][1], o[2 * jj + 1][2], o[2 * jj + 1][3], a0l, a2l, v10, v11);
      }
      if (last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the new token's row before a later copy overwrites it
      __syncwarp();  // every lane is done reading this stage before lane 0 may refill it (next iteration's issue)
      if (++c_s == STAGES) { c_s = 0; c_ph ^= 1; }
    }
#pragma unroll
    for (int x = 1; x < 32; x <<= 1) l += __shfl_xor_sync(0xffffffffu, l, x);
    const float inv = 1.0f / l;
    // sum the 8 rows and undo the column skew: at step b a lane keeps the accumulators j with bit b clear and adds the
    // partner row's accumulator j ^ (1 << b) (the same dims column there); lane (g, t) ends with dims column g in o[0]
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      o[j][0] += __shfl_xor_sync(0xffffffffu, o[j + 1][0], 4);
      o[j][1] += __shfl_xor_sync(0xffffffffu, o[j + 1][1], 4);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      o[j][0] += __shfl_xor_sync(0xffffffffu, o[j + 2][0], 8);
      o[j][1] += __shfl_xor_sync(0xffffffffu, o[j + 2][1], 8);
    }
    o[0][0] += __shfl_xor_sync(0xffffffffu, o[4][0], 16);
    o[0][1] += __shfl_xor_sync(0xffffffffu, o[4][1], 16);
    // out row of the item = item * 64 elements (row * d + head * 64): dims 8 g + 2 t, + 1 -> one 128-byte row per warp
    const float y0 = o[0][0] * inv, y1 = o[0][1] * inv;
    const uint32_t yh = pack_bf16x2(y0, y1);
    *reinterpret_cast<uint32_t*>(out + (size_t)item * HD + 8 * g + 2 * t) = yh;
    if (F16) {  // the bf16x2 engine's operand: remainder of the bf16 rounding
      const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yh));
      *reinterpret_cast<uint32_t*>(out_lo + (size_t)item * HD + 8 * g + 2 * t) = pack_bf16x2(y0 - yf.x, y1 - yf.y);
    }
  }
  trace_end(step_trace, tslot);
}

// ---------------------------------------------------------------------------------------------------------------
// Beam-search decode attention with the image prefix read ONCE per image.  The `beams` hypotheses of an image share the
// first n_prefix cached positions (the prefill row of the image, see INDIRECT above): at config 3 (prefix 40, 5 beams, <= 30 generated
// tokens) those are 58-100 % of the keys every hypothesis walks, so one warp takes an (image, head) item and runs the beams as the ROWS of
// the m16n8k16 tile: a prefix chunk (16 keys, contiguous) is scored against all beams at once, a generated chunk (16 positions of ONE
// hypothesis, each from the cache row its ancestry table names) only counts for that hypothesis' row (the other rows are masked).
// Per-row online softmax (FlashAttention-2 register reuse: the S accumulators become the A operand of P.V, split hi + lo).
// Chunks land in a per-warp cp.async ring with the 16-byte columns XOR-swizzled by the key index, so ldmatrix (K) / ldmatrix.trans (V) read
// true fragments without bank conflicts.  The new token of each hypothesis goes into the free slot of its last chunk and is appended to
// its own cache row.  HF semantics unchanged: softmax over [prefix | own generated history | new token] (HF:models/gpt2/modeling_gpt2.py:
// 144-226 with DynamicCache.reorder_cache replaced by the ancestry table).
// ---------------------------------------------------------------------------------------------------------------
template <int WARPS, int STAGES> struct BeamAttnCfg { static constexpr int SMEM_BYTES = WARPS * STAGES * MMA_STAGE_BYTES + 128; };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <bool F16, int BEAM_ATTN_WARPS, int BEAM_ATTN_STAGES>
__global__ void __launch_bounds__(BEAM_ATTN_WARPS * 32, 1) attn_decode_beam_kernel(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, bf16* out_lo,
                                                                                   const int* d_pos, int images, int nb, int H, int t_max, const int* anc,
                                                                                   int anc_ld, int n_prefix, StepTrace step_trace) {
  constexpr int STAGES = BEAM_ATTN_STAGES;
  extern __shared__ uint8_t dec_smem_raw[];
  const uint32_t smem_base = (dec_smem_u32(dec_smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_base + warp * (STAGES * MMA_STAGE_BYTES);
  pdl_launch_dependents();
  // zero the ring: slots no copy has filled are multiplied by P = 0 and must not hold NaN / Inf
  for (int i = lane; i < STAGES * MMA_STAGE_BYTES / 16; i += 32)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ring + i * 16), "r"(0u) : "memory");
  __syncwarp();
  pdl_wait();
  const int tslot = trace_begin(step_trace, TRACE_ATTN_DECODE, 1);
  const int pos = __ldcg(d_pos);       // positions already cached == position of the new token
  const int ngen = pos - n_prefix;     // cached generated positions of every hypothesis
  const int d = H * HD;
  const int n_items = images * H;
  const int wstride = gridDim.x * BEAM_ATTN_WARPS;
  const int w0 = blockIdx.x * BEAM_ATTN_WARPS + warp;
  const int nchp = (n_prefix + MMA_CHUNK - 1) / MMA_CHUNK;  // prefix chunks (shared by the beams)
  const int nchg = ngen / MMA_CHUNK + 1;                    // chunks of one hypothesis' generated keys + its new token
  const int nch = nchp + nb * nchg;
  const int my_items = w0 < n_items ? (n_items - 1 - w0) / wstride + 1 : 0;
  const int total_chunks = my_items * nch;
  const int g = lane >> 2, t = lane & 3;
  // this lane's copy pieces of a chunk: keys (lane >> 3) + 4 i, 16-byte column lane & 7 -> swizzled column (lane & 7) ^ (key & 7).
  // (every offset below is a 32-bit element index: the launcher checks that a cache plane has fewer than 2^31 elements)
  const int pj = lane >> 3, pc = lane & 7;
  const uint32_t hstride = (uint32_t)t_max * HD;  // elements between two (row, head) planes
  uint32_t sw_off[4];                             // swizzled byte offset of piece i inside the K (or V) half of a stage
#pragma unroll
  for (int i = 0; i < 4; ++i) sw_off[i] = (uint32_t)((pj + 4 * i) * (HD * 2) + ((pc ^ ((pj + 4 * i) & 7)) << 4));

  // producer cursor (no divisions on the per-chunk path): item -> (image, head), chunk -> prefix chunk p_c < nchp or (beam p_b, chunk p_g)
  int p_item = w0, p_img = w0 / H, p_h = w0 - (w0 / H) * H, p_c = 0, p_b = 0, p_g = 0, p_s = 0;
  // cache rows of the generated keys this lane copies in the NEXT chunk to be issued, looked up one issue ahead (an L2 round trip
  // that would otherwise sit in front of every generated chunk)
  int src_next[4];
  auto lookup = [&]() {
    if (p_c < nchp) return;
    const int* arow = anc + (size_t)(p_img * nb + p_b) * anc_ld + p_g * MMA_CHUNK + pj;
#pragma unroll
    for (int i = 0; i < 4; ++i) src_next[i] = p_g * MMA_CHUNK + pj + 4 * i < ngen ? __ldcg(arow + 4 * i) : -1;
  };
  auto issue_next = [&]() {
    const uint32_t dst = ring + p_s * MMA_STAGE_BYTES;
    if (p_c < nchp) {
      const int k0 = p_c * MMA_CHUNK, nkeys = min(MMA_CHUNK, n_prefix - k0);
      const uint32_t base = ((uint32_t)(p_img * nb) * H + p_h) * hstride + (uint32_t)(k0 + pj) * HD + pc * 8;  // the image's prefill row
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (pj + 4 * i < nkeys) {
          cp_async16(dst + sw_off[i], kcache + base + i * (4 * HD));
          cp_async16(dst + MMA_CHUNK * HD * 2 + sw_off[i], vcache + base + i * (4 * HD));
        }
    } else {
      const uint32_t rel = (uint32_t)p_h * hstride + (uint32_t)(n_prefix + p_g * MMA_CHUNK + pj) * HD + pc * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (src_next[i] >= 0) {
          const uint32_t off = (uint32_t)src_next[i] * (H * hstride) + rel + i * (4 * HD);
          cp_async16(dst + sw_off[i], kcache + off);
          cp_async16(dst + MMA_CHUNK * HD * 2 + sw_off[i], vcache + off);
        }
    }
    // advance
    if (++p_s == STAGES) p_s = 0;
    if (++p_c > nchp && ++p_g == nchg) { p_g = 0; ++p_b; }
    if (p_c == nchp) { p_b = 0; p_g = 0; }
    if (p_c == nch) {
      p_c = 0; p_b = 0; p_g = 0;
      p_item += wstride;
      p_img = p_item / H;
      p_h = p_item - p_img * H;
    }
    if (p_item < n_items) lookup();
  };
  int issued = 0;
  if (my_items > 0) lookup();
  for (; issued < STAGES - 1; ++issued) {
    if (issued < total_chunks) issue_next();
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // q of the item's hypotheses as A fragments (row g = beam g; rows >= nb zero), requested one item ahead
  uint32_t qa_n[8];
  auto load_q = [&](int img, int h) {
#pragma unroll
    for (int j = 0; j < 8; ++j) qa_n[j] = 0u;
    if (g < nb) {
      const uint32_t* q32 = reinterpret_cast<const uint32_t*>(qkv + (size_t)(img * nb + g) * 3 * d + h * HD);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        qa_n[2 * ks] = __ldcg(q32 + 8 * ks + t);          // dims 16 ks + 2 t, + 1
        qa_n[2 * ks + 1] = __ldcg(q32 + 8 * ks + 4 + t);  // dims 16 ks + 8 + 2 t, + 1
      }
    }
  };
  int c_img = w0 / H, c_h = w0 - (w0 / H) * H;
  if (my_items > 0) load_q(c_img, c_h);
  const int r8 = lane & 7, mlo = (lane >> 3) & 1, mhi = lane >> 4;
  const uint32_t k_row_off = (uint32_t)((8 * mhi + r8) * (HD * 2));
  const uint32_t v_row_off = (uint32_t)(MMA_CHUNK * HD * 2 + (8 * mlo + r8) * (HD * 2));
  int c_s = 0;
  for (int ii = 0; ii < my_items; ++ii) {
    const int img = c_img, h = c_h;
    uint32_t qa[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qa[i] = scale_eighth<F16>(qa_n[i]);
    if (ii + 1 < my_items) {
      const int nitem = w0 + (ii + 1) * wstride;
      c_img = nitem / H;
      c_h = nitem - c_img * H;
      load_q(c_img, c_h);
    }
    // the hypotheses' new k / v rows: 16 pieces of 16 bytes per beam (8 of k, 8 of v), piece lane + 32 i
    uint4 newkv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pi = lane + 32 * i, b = pi >> 4;
      if (b < nb) newkv[i] = __ldcg(reinterpret_cast<const uint4*>(qkv + (size_t)(img * nb + b) * 3 * d + d + ((pi >> 3) & 1) * d + h * HD) + (pi & 7));
    }
    float m = -INFINITY, l = 0.f;  // row g: running maximum (uniform over the quad) and this lane's share of the running sum
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }
    int beam = -1, gch = 0;  // consumer cursor: prefix chunks (beam < 0), then chunk gch of hypothesis `beam`
    for (int c = 0; c < nch; ++c) {
      if (issued < total_chunks) issue_next();
      asm volatile("cp.async.commit_group;" ::: "memory");
      ++issued;
      const uint32_t st = ring + c_s * MMA_STAGE_BYTES;
      if (c == nchp) { beam = 0; gch = 0; }
      // which rows this chunk counts for, how many key slots it fills
      int nkeys;
      if (beam < 0) {
        nkeys = min(MMA_CHUNK, n_prefix - c * MMA_CHUNK);
      } else {
        const int cached = min(MMA_CHUNK, ngen - gch * MMA_CHUNK);
        nkeys = cached;
        if (gch == nchg - 1) {  // the hypothesis' new token: slot `cached` (< 16) of its last chunk, and its own cache row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pi = lane + 32 * i;
            if ((pi >> 4) == beam) {
              const int kv = (pi >> 3) & 1, cc = pi & 7;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + kv * (MMA_CHUNK * HD * 2) + cached * (HD * 2) + ((cc ^ (cached & 7)) << 4)),
                           "r"(newkv[i].x), "r"(newkv[i].y), "r"(newkv[i].z), "r"(newkv[i].w) : "memory");
              const uint32_t coff = ((uint32_t)(img * nb + beam) * H + h) * hstride + (uint32_t)pos * HD + cc * 8;
              *reinterpret_cast<uint4*>((kv ? vcache : kcache) + coff) = newkv[i];  // append (HF:cache_utils.py:102-121)
            }
          }
          nkeys = cached + 1;
        }
      }
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
      __syncwarp();  // every lane's pieces of this chunk (and the new token's row) are visible
      // ---- S = Q . K^T: rows = beams, 16 key slots ----
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b00, b01, b10, b11;
        ldsm_x4(st + k_row_off + 16 * ((2 * ks + mlo) ^ r8), b00, b01, b10, b11);
        mma_16816_top<F16>(s0[0], s0[1], s0[2], s0[3], qa[2 * ks], qa[2 * ks + 1], b00, b01);
        mma_16816_top<F16>(s1[0], s1[1], s1[2], s1[3], qa[2 * ks], qa[2 * ks + 1], b10, b11);
      }
      const bool row_on = g < nb && (beam < 0 || beam == g);
      const float x0 = (row_on && 2 * t < nkeys) ? s0[0] : -INFINITY, x1 = (row_on && 2 * t + 1 < nkeys) ? s0[1] : -INFINITY;
      const float x2 = (row_on && 8 + 2 * t < nkeys) ? s1[0] : -INFINITY, x3 = (row_on && 9 + 2 * t < nkeys) ? s1[1] : -INFINITY;
      float cm = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
      const float mn = fmaxf(m, cm);
      const float mref = (mn == -INFINITY) ? 0.f : mn * LOG2E;  // rows this chunk (and all before it) do not count for: everything stays 0
      const float scale = exp2f(m * LOG2E - mref);              // 0 while m = -inf
      m = mn;
      const float p0 = exp2f(fmaf(x0, LOG2E, -mref)), p1 = exp2f(fmaf(x1, LOG2E, -mref)), p2 = exp2f(fmaf(x2, LOG2E, -mref)), p3 = exp2f(fmaf(x3, LOG2E, -mref));
      l = l * scale + ((p0 + p1) + (p2 + p3));
      const float p0h = round_op<F16>(p0), p1h = round_op<F16>(p1), p2h = round_op<F16>(p2), p3h = round_op<F16>(p3);
      const uint32_t a0h = pack_op2<F16>(p0h, p1h), a2h = pack_op2<F16>(p2h, p3h);
      const uint32_t a0l = pack_op2<F16>(p0 - p0h, p1 - p1h), a2l = pack_op2<F16>(p2 - p2h, p3 - p3h);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j][0] *= scale; o[j][1] *= scale; }
      // ---- O += P . V ----
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t v00, v01, v10, v11;
        ldsm_x4_trans(st + v_row_off + 16 * ((2 * jj + mhi) ^ r8), v00, v01, v10, v11);
        mma_16816_top<F16>(o[2 * jj][0], o[2 * jj][1], o[2 * jj][2], o[2 * jj][3], a0h, a2h, v00, v01);
        mma_16816_top<F16>(o[2 * jj + 1][0], o[2 * jj + 1][1], o[2 * jj + 1][2], o[2 * jj + 1][3], a0h, a2h, v10, v11);
        mma_16816_top<F16>(o[2 * jj][0], o[2 * jj][1], o[2 * jj][2], o[2 * jj][3], a0l, a2l, v00, v01);
        mma_16816_top<F16>(o[2 * jj + 1][0], o[2 * jj + 1][1], o[2 * jj + 1][2], o[2 * jj + 1][3], a0l, a2l, v10, v11);
      }
      __syncwarp();  // every lane is done reading this stage before the next iteration's copies may overwrite it
      if (++c_s == STAGES) c_s = 0;
      if (beam >= 0 && ++gch == nchg) { gch = 0; ++beam; }
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (g < nb) {
      const float inv = 1.0f / l;
      bf16* orow = out + (size_t)(img * nb + g) * d + h * HD + 2 * t;
      bf16* lrow = F16 ? out_lo + (size_t)(img * nb + g) * d + h * HD + 2 * t : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y0 = o[j][0] * inv, y1 = o[j][1] * inv;
        const uint32_t yh = pack_bf16x2(y0, y1);
        *reinterpret_cast<uint32_t*>(orow + 8 * j) = yh;
        if (F16) {
          const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yh));
          *reinterpret_cast<uint32_t*>(lrow + 8 * j) = pack_bf16x2(y0 - yf.x, y1 - yf.y);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  trace_end(step_trace, tslot);
}

#define GIC_BEAM_ATTN_VARIANTS(X) X(8, 4) X(12, 4) X(16, 3)
static int g_dec_sms = 0;
constexpr int DEC_PRODUCT_VARIANT = 12;  // mma.sync kernel, 12 warps x 4 stages (profiles/r1aa_microbench.txt)
static int g_dec_variant = DEC_PRODUCT_VARIANT;  // microbenchmark / test knob (attn_decode_set_variant); < 0 = back to the product shape
void attn_decode_set_variant(int v) { g_dec_variant = v < 0 ? DEC_PRODUCT_VARIANT : v; }
#define GIC_DEC_VARIANTS(X) X(0, 16, 16, 3) X(1, 32, 8, 3)
#define GIC_DEC_MMA_VARIANTS(X) X(10, 16, 3) X(11, 8, 6) X(12, 12, 4) X(13, 8, 5) X(14, 4, 12)
// opt the bulk-copy kernels into their shared memory size once (engine creation: outside any stream capture)
int attn_decode_configure() {
  if (g_dec_sms > 0) return GIC_OK;
#define X(ID, W, C, S) \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_bulk_kernel<W, C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg<W, C, S>::SMEM_BYTES));
  GIC_DEC_VARIANTS(X)
#undef X
#define X(ID, W, S) \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_mma_kernel<W, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecMmaCfg<W, S>::SMEM_BYTES));
  GIC_DEC_MMA_VARIANTS(X)
#undef X
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_mma_kernel<12, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecMmaCfg<12, 4>::SMEM_BYTES));
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_mma_kernel<12, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecMmaCfg<12, 4>::SMEM_BYTES));
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_mma_kernel<12, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecMmaCfg<12, 4>::SMEM_BYTES));
#define X(W, S)                                                                                                                                           \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_beam_kernel<false, W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, BeamAttnCfg<W, S>::SMEM_BYTES)); \
  GIC_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_beam_kernel<true, W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, BeamAttnCfg<W, S>::SMEM_BYTES));
  GIC_BEAM_ATTN_VARIANTS(X)
#undef X
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  g_dec_sms = sms;
  return GIC_OK;
}

static int launch_attn_decode_bulk(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, const int* d_pos, int rows, int H, int t_max,
                                   cudaStream_t st, const int* row_map) {
  GIC_TRY(attn_decode_configure());
  GIC_REQUIRE(row_map == nullptr || g_dec_variant >= 10, "attn_decode: the SIMT comparison kernels do not take a row map");
  const int sms = cta_limit() > 0 && cta_limit() < g_dec_sms ? cta_limit() : g_dec_sms;
  const int items = rows * H;
#define X(ID, W, C, S)                                                                                                              \
  if (g_dec_variant == ID) {                                                                                                         \
    const int grid = min(sms, ceil_div(items, W));                                                                                   \
    GIC_CHECK_CUDA(launch_kernel(attn_decode_bulk_kernel<W, C, S>, dim3(grid), dim3(W * 32), (size_t)DecCfg<W, C, S>::SMEM_BYTES, st, qkv, kcache, \
                                 vcache, out, d_pos, rows, H, t_max));                                                               \
    note_launch();                                                                                                                   \
    return GIC_OK;                                                                                                                   \
  }
  GIC_DEC_VARIANTS(X)
#undef X
#define X(ID, W, S)                                                                                                                  \
  if (g_dec_variant == ID) {                                                                                                         \
    const int grid = min(sms, ceil_div(items, W));                                                                                   \
    GIC_CHECK_CUDA(launch_kernel(attn_decode_mma_kernel<W, S, false>, dim3(grid), dim3(W * 32), (size_t)DecMmaCfg<W, S>::SMEM_BYTES, st, qkv, kcache, \
                                 vcache, out, d_pos, rows, H, t_max, (const int*)nullptr, 0, 0, 1, trace_desc(), (bf16*)nullptr, row_map));        \
    note_launch();                                                                                                                   \
    return GIC_OK;                                                                                                                   \
  }
  GIC_DEC_MMA_VARIANTS(X)
#undef X
  set_error("attn_decode: unknown kernel variant %d", g_dec_variant);
  return GIC_ERR_UNSUPPORTED;
}

// GIC_BEAM_SHARED_PREFIX=0: every hypothesis walks its whole context on its own (attn_decode_mma_kernel INDIRECT; measurement / tests)
constexpr int BEAM_ATTN_DEFAULT_WARPS = 12;  // c3 decode step, bf16: 140 / 124.5 / 124.3 us per launch with 8 / 12 / 16 warps (issue bound; 16 spills) -- profiles/r2u_beam_attn_warps.txt
static int g_beam_shared_override = -1;  // test knob (attn_decode_set_beam_shared): 0 / 1 force the kernel, < 0 = the environment's choice
void attn_decode_set_beam_shared(int v) { g_beam_shared_override = v; }
static bool beam_shared_prefix_enabled() {
  if (g_beam_shared_override >= 0) return g_beam_shared_override != 0;
  const char* v = getenv("GIC_BEAM_SHARED_PREFIX");  // (read per launch: the tests switch it inside one process)
  return !(v && v[0] == '0');
}

// beam search without cache reordering (see attn_decode_mma_kernel INDIRECT and attn_decode_beam_kernel)
int launch_attn_decode_indirect(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, const int* d_pos, int rows, int H, int t_max, const int* anc,
                                int anc_ld, int n_prefix, int beams, cudaStream_t st, bf16* out_lo) {
  GIC_TRY(attn_decode_configure());
  GIC_REQUIRE(anc != nullptr && beams >= 1 && n_prefix >= 0, "attn_decode_indirect: bad ancestry arguments");
  const int sms = cta_limit() > 0 && cta_limit() < g_dec_sms ? cta_limit() : g_dec_sms;
  if (beam_shared_prefix_enabled() && beams >= 2 && beams <= 8 && rows % beams == 0 && (size_t)rows * H * t_max * HD < ((size_t)1 << 31)) {
    // the beams of an image as the rows of one MMA tile: the shared prefix is read once per image (attn_decode_beam_kernel)
    const int images = rows / beams;
    const char* wv = getenv("GIC_BEAM_ATTN_WARPS");  // measurement knob: 8 | 12 | 16 warps per CTA
    const int want = wv ? atoi(wv) : BEAM_ATTN_DEFAULT_WARPS;
#define X(W, S)                                                                                                                                                \
    if (want == W) {                                                                                                                                             \
      const int bgrid = min(sms, ceil_div(images * H, W));                                                                                                       \
      if (out_lo)                                                                                                                                                \
        GIC_CHECK_CUDA(launch_kernel(attn_decode_beam_kernel<true, W, S>, dim3(bgrid), dim3(W * 32), (size_t)BeamAttnCfg<W, S>::SMEM_BYTES, st, qkv, kcache, vcache, out, \
                                     out_lo, d_pos, images, beams, H, t_max, anc, anc_ld, n_prefix, trace_desc()));                                            \
      else                                                                                                                                                       \
        GIC_CHECK_CUDA(launch_kernel(attn_decode_beam_kernel<false, W, S>, dim3(bgrid), dim3(W * 32), (size_t)BeamAttnCfg<W, S>::SMEM_BYTES, st, qkv, kcache, vcache, out, \
                                     (bf16*)nullptr, d_pos, images, beams, H, t_max, anc, anc_ld, n_prefix, trace_desc()));                                    \
      note_launch();                                                                                                                                             \
      return GIC_OK;                                                                                                                                             \
    }
    GIC_BEAM_ATTN_VARIANTS(X)
#undef X
    set_error("attn_decode_indirect: GIC_BEAM_ATTN_WARPS=%d is not one of 8, 12, 16", want);
    return GIC_ERR_UNSUPPORTED;
  }
  const int grid = min(sms, ceil_div(rows * H, 12));
  if (out_lo)  // fp16 q / k / v / cache, hi + lo output (bf16x2 engine)
    GIC_CHECK_CUDA(launch_kernel(attn_decode_mma_kernel<12, 4, true, true>, dim3(grid), dim3(12 * 32), (size_t)DecMmaCfg<12, 4>::SMEM_BYTES, st, qkv, kcache,
                                 vcache, out, d_pos, rows, H, t_max, anc, anc_ld, n_prefix, beams, trace_desc(), out_lo, (const int*)nullptr));
  else
    GIC_CHECK_CUDA(launch_kernel(attn_decode_mma_kernel<12, 4, true>, dim3(grid), dim3(12 * 32), (size_t)DecMmaCfg<12, 4>::SMEM_BYTES, st, qkv, kcache, vcache,
                                 out, d_pos, rows, H, t_max, anc, anc_ld, n_prefix, beams, trace_desc(), (bf16*)nullptr, (const int*)nullptr));
  note_launch();
  return GIC_OK;
}

// decode attention of the bf16x2 engine: fp16 q | k | v and KV cache (2-byte elements, typed bf16* for the shared plumbing), output as a
// bf16 hi + lo pair
int launch_attn_decode_f16(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out_hi, bf16* out_lo, const int* d_pos, int rows, int H, int t_max,
                           cudaStream_t st, const int* row_map) {
  GIC_TRY(attn_decode_configure());
  GIC_REQUIRE(out_hi && out_lo, "attn_decode_f16: needs both output halves");
  const int sms = cta_limit() > 0 && cta_limit() < g_dec_sms ? cta_limit() : g_dec_sms;
  const int grid = min(sms, ceil_div(rows * H, 12));
  GIC_CHECK_CUDA(launch_kernel(attn_decode_mma_kernel<12, 4, false, true>, dim3(grid), dim3(12 * 32), (size_t)DecMmaCfg<12, 4>::SMEM_BYTES, st, qkv, kcache,
                               vcache, out_hi, d_pos, rows, H, t_max, (const int*)nullptr, 0, 0, 1, trace_desc(), out_lo, row_map));
  note_launch();
  return GIC_OK;
}
bool attn_decode_indirect_available() {
  const char* v = getenv("GIC_ATTN_SIMPLE");
  const char* r = getenv("GIC_BEAM_REORDER");  // =1: physical cache reorder (kv_reorder_kernel) instead of the ancestry table
  return !(v && v[0] == '1') && !(r && r[0] == '1');
}

static bool decode_bulk_enabled() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("GIC_ATTN_SIMPLE"); on = (v && v[0] == '1') ? 0 : 1; }
  return on == 1;
}
template <typename T> struct IsBf16 { static constexpr bool value = false; };
template <> struct IsBf16<bf16> { static constexpr bool value = true; };

template <typename T>
int launch_attn_decode(const T* qkv, T* kcache, T* vcache, ActOut out, const int* d_pos, int rows, int H, int t_max, cudaStream_t st, const int* row_map) {
  if (IsBf16<T>::value && out.hi && !out.lo && !out.f32 && decode_bulk_enabled())
    return launch_attn_decode_bulk((const bf16*)qkv, (bf16*)kcache, (bf16*)vcache, out.hi, d_pos, rows, H, t_max, st, row_map);
  const int warps = rows * H;
  const int blocks = ceil_div(warps, 4);
  const size_t smem = 0;
  GIC_CHECK_CUDA(launch_kernel(attn_decode_kernel<T>, dim3(blocks), dim3(128), smem, st, qkv, kcache, vcache, out, d_pos, rows, H, t_max, row_map));
  note_launch();
  return GIC_OK;
}
template int launch_attn_decode<float>(const float*, float*, float*, ActOut, const int*, int, int, int, cudaStream_t, const int*);
template int launch_attn_decode<bf16>(const bf16*, bf16*, bf16*, ActOut, const int*, int, int, int, cudaStream_t, const int*);

// ---------------------------------------------------------------------------------------------------------------
// Sequence attention: one block per (row, head), K/V of the row staged in shared memory as fp32, one warp per query.
//   CAUSAL + cache write  -> GPT-2 prefill over the prefix tokens
//   bidirectional          -> transformer-mapper encoder layers
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int HDIM, bool CAUSAL>
__global__ void __launch_bounds__(128) attn_seq_kernel(const T* qkv, T* kcache, T* vcache, ActOut out, int S, int H,
                                                       int t_max, int cache_row_mult, float scale) {
  constexpr int DPL = (HDIM + 31) / 32;  // dims per lane (lanes >= HDIM idle when HDIM < 32)
  extern __shared__ float sm[];
  float* Ks = sm;                       // [S][HDIM]
  float* Vs = Ks + (size_t)S * HDIM;    // [S][HDIM]
  float* Sc = Vs + (size_t)S * HDIM;    // [4 warps][S]
  const int row = blockIdx.x / H, h = blockIdx.x % H;
  const int d = H * HDIM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* base = qkv + (size_t)row * S * 3 * d + h * HDIM;
  pdl_launch_dependents();
  pdl_wait();

  for (int i = threadIdx.x; i < S * HDIM; i += blockDim.x) {
    const int t = i / HDIM, c = i % HDIM;
    const T kv = base[(size_t)t * 3 * d + d + c];
    const T vv = base[(size_t)t * 3 * d + 2 * d + c];
    Ks[i] = to_f32(kv);
    Vs[i] = to_f32(vv);
    if (kcache) {
      const size_t ci = (((size_t)row * cache_row_mult * H + h) * t_max + t) * HDIM + c;
      kcache[ci] = kv;
      vcache[ci] = vv;
    }
  }
  __syncthreads();

  float* sc = Sc + (size_t)warp * S;
  for (int t = warp; t < S; t += 4) {
    float q[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) q[i] = (lane + 32 * i < HDIM) ? to_f32(base[(size_t)t * 3 * d + lane + 32 * i]) * scale : 0.f;
    const int n = CAUSAL ? t + 1 : S;
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (lane + 32 * i < HDIM) s = fmaf(q[i], Ks[j * HDIM + lane + 32 * i], s);
      s = warp_sum(s);
      if (lane == 0) sc[j] = s;
      mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float p = expf(sc[j] - mx);
      sc[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.f;
    for (int j = 0; j < n; ++j) {
      const float p = sc[j];
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (lane + 32 * i < HDIM) o[i] = fmaf(p, Vs[j * HDIM + lane + 32 * i], o[i]);
    }
    const float inv = 1.0f / sum;
    const size_t o0 = ((size_t)row * S + t) * d + h * HDIM;
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      if (lane + 32 * i < HDIM) out.write(o0 + lane + 32 * i, o[i] * inv);
    __syncwarp();
  }
}

template <typename T, int HDIM, bool CAUSAL>
static int launch_attn_seq(const T* qkv, T* kcache, T* vcache, ActOut out, int B, int S, int H, int t_max, int cache_row_mult,
                           cudaStream_t st) {
  const size_t smem = ((size_t)2 * S * HDIM + 4 * S) * sizeof(float);
  GIC_REQUIRE(smem <= 200 * 1024, "attention: sequence %d x head_dim %d does not fit in shared memory", S, HDIM);
  auto kern = attn_seq_kernel<T, HDIM, CAUSAL>;
  if (smem > 48 * 1024) GIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GIC_CHECK_CUDA(launch_kernel(kern, dim3(B * H), dim3(128), smem, st, qkv, kcache, vcache, out, S, H, t_max, cache_row_mult,
                               1.0f / sqrtf((float)HDIM)));
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 causal prefill attention on mma.sync (head_dim 64, up to 64 prefix tokens): one warp per (row, head).  q / k / v of
// the row's S tokens are staged once in shared memory with the 16-byte columns XOR-swizzled by the row (conflict-free
// ldmatrix), K / V go to the cache on the way in, and the S x S problem runs as 16 x 16 blocks with an online softmax
// (FlashAttention-2 register reuse: the score accumulators become the P operand; P as bf16 hi + lo like the decode kernel).
// attn_seq_kernel does the same job in 68 us per layer for B = 1024, P = 10 (scalar loads, one shuffle tree per key).
// ---------------------------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// F16: fp16 q / k / v / cache and a bf16 hi + lo output pair (out, out_lo), as in attn_decode_mma_kernel
template <int NBLK, bool F16 = false>  // 16-token blocks staged per item: S <= 16 NBLK
__global__ void __launch_bounds__(128) attn_prefill_mma_kernel(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, int n_items, int S, int H,
                                                             int t_max, int cache_row_mult, bf16* out_lo) {
  constexpr int SP = 16 * NBLK;             // staged rows per matrix
  constexpr int MAT_BYTES = SP * HD * 2;    // q, k or v of one item
  extern __shared__ uint8_t pre_smem_raw[];
  const uint32_t smem_base = (dec_smem_u32(pre_smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t qs = smem_base + warp * (3 * MAT_BYTES), ks = qs + MAT_BYTES, vs = ks + MAT_BYTES;
  const int item = blockIdx.x * 4 + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (item >= n_items) return;
  const int row = item / H, h = item - row * H;
  const int d = H * HD;
  // ---- stage q (pre-scaled by 1/sqrt(64)), k, v; append k, v to the cache (HF:cache_utils.py:102-121) ----
  for (int i = lane; i < SP * 8; i += 32) {
    const int t = i >> 3, c = i & 7;
    uint4 qv = make_uint4(0, 0, 0, 0), kv = qv, vv = qv;  // rows >= S: zeros (P = 0 times V must stay finite)
    if (t < S) {
      const bf16* src = qkv + ((size_t)row * S + t) * 3 * d + h * HD + c * 8;
      qv = *reinterpret_cast<const uint4*>(src);
      kv = *reinterpret_cast<const uint4*>(src + d);
      vv = *reinterpret_cast<const uint4*>(src + 2 * d);
      const size_t ci = (((size_t)row * cache_row_mult * H + h) * t_max + t) * HD + c * 8;
      *reinterpret_cast<uint4*>(kcache + ci) = kv;
      *reinterpret_cast<uint4*>(vcache + ci) = vv;
      qv.x = scale_eighth<F16>(qv.x); qv.y = scale_eighth<F16>(qv.y); qv.z = scale_eighth<F16>(qv.z); qv.w = scale_eighth<F16>(qv.w);
    }
    const uint32_t off = (uint32_t)(t * (HD * 2) + ((c ^ (t & 7)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(qs + off), "r"(qv.x), "r"(qv.y), "r"(qv.z), "r"(qv.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ks + off), "r"(kv.x), "r"(kv.y), "r"(kv.z), "r"(kv.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(vs + off), "r"(vv.x), "r"(vv.y), "r"(vv.z), "r"(vv.w) : "memory");
  }
  __syncwarp();
  const int g = lane >> 2, t4 = lane & 3;
  const int r8 = lane & 7, mlo = (lane >> 3) & 1, mhi = lane >> 4;
  const int nblk = (S + 15) >> 4;
  for (int qb = 0; qb < nblk; ++qb) {
    // Q fragments of the 16 queries: matrix m of the x4 = rows 8 (m & 1) + r, 16-byte column 2 ks + (m >> 1)
    uint32_t qa[4][4];
    {
      const int qrow = qb * 16 + 8 * mlo + r8;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ldsm_x4(qs + qrow * (HD * 2) + (((2 * kk + mhi) ^ (qrow & 7)) << 4), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g + 8 of the block
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }
    for (int kb = 0; kb <= qb; ++kb) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int krow = kb * 16 + 8 * mhi + r8;  // K: keys 8 (m >> 1) + r, column 2 ks + (m & 1)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t b00, b01, b10, b11;
          ldsm_x4(ks + krow * (HD * 2) + (((2 * kk + mlo) ^ (krow & 7)) << 4), b00, b01, b10, b11);
          mma_16816<F16>(s0, qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b00, b01);
          mma_16816<F16>(s1, qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b10, b11);
        }
      }
      // causal mask (HF GPT2Attention is_causal): key index <= query index; s?[0..1] row g, s?[2..3] row g + 8
      const int q0 = qb * 16 + g, q1 = q0 + 8, kbase = kb * 16 + 2 * t4;
      if (kbase > q0) s0[0] = -INFINITY;
      if (kbase + 1 > q0) s0[1] = -INFINITY;
      if (kbase + 8 > q0) s1[0] = -INFINITY;
      if (kbase + 9 > q0) s1[1] = -INFINITY;
      if (kbase > q1) s0[2] = -INFINITY;
      if (kbase + 1 > q1) s0[3] = -INFINITY;
      if (kbase + 8 > q1) s1[2] = -INFINITY;
      if (kbase + 9 > q1) s1[3] = -INFINITY;
      float c0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), c1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
      // (block kb = 0 always holds key 0 <= every query, so the running maxima are finite from the first block on)
      const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
      const float sc0 = exp2f((m0 - n0) * LOG2E), sc1 = exp2f((m1 - n1) * LOG2E);
      m0 = n0; m1 = n1;
      const float nl0 = n0 * LOG2E, nl1 = n1 * LOG2E;
      float p[8];
      p[0] = exp2f(fmaf(s0[0], LOG2E, -nl0)); p[1] = exp2f(fmaf(s0[1], LOG2E, -nl0));
      p[2] = exp2f(fmaf(s0[2], LOG2E, -nl1)); p[3] = exp2f(fmaf(s0[3], LOG2E, -nl1));
      p[4] = exp2f(fmaf(s1[0], LOG2E, -nl0)); p[5] = exp2f(fmaf(s1[1], LOG2E, -nl0));
      p[6] = exp2f(fmaf(s1[2], LOG2E, -nl1)); p[7] = exp2f(fmaf(s1[3], LOG2E, -nl1));
      l0 = l0 * sc0 + ((p[0] + p[1]) + (p[4] + p[5]));
      l1 = l1 * sc1 + ((p[2] + p[3]) + (p[6] + p[7]));
      uint32_t ph[4], pl[4];  // A fragments of P: (row g, keys 2t..), (row g + 8, keys 2t..), (row g, keys 8 + 2t..), (row g + 8, keys 8 + 2t..)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float h0 = round_op<F16>(p[2 * u]), h1 = round_op<F16>(p[2 * u + 1]);
        ph[u] = pack_op2<F16>(h0, h1);
        pl[u] = pack_op2<F16>(p[2 * u] - h0, p[2 * u + 1] - h1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1; }
      {
        const int vrow = kb * 16 + 8 * mlo + r8;  // V (trans): keys 8 (m & 1) + r, column 2 jj + (m >> 1)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t v00, v01, v10, v11;
          ldsm_x4_trans(vs + vrow * (HD * 2) + (((2 * jj + mhi) ^ (vrow & 7)) << 4), v00, v01, v10, v11);
          mma_16816<F16>(o[2 * jj], ph[0], ph[1], ph[2], ph[3], v00, v01);
          mma_16816<F16>(o[2 * jj + 1], ph[0], ph[1], ph[2], ph[3], v10, v11);
          mma_16816<F16>(o[2 * jj], pl[0], pl[1], pl[2], pl[3], v00, v01);
          mma_16816<F16>(o[2 * jj + 1], pl[0], pl[1], pl[2], pl[3], v10, v11);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // the block's 16 output rows go back through the (now consumed) q rows of this block, then out as whole 16-byte chunks
    // (F16: a second pass carries the remainders of the bf16 rounding to out_lo)
#pragma unroll 1
    for (int pass = 0; pass < (F16 ? 2 : 1); ++pass) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ra = qb * 16 + g, rb = ra + 8;
        float y0 = o[j][0] * i0, y1 = o[j][1] * i0, y2 = o[j][2] * i1, y3 = o[j][3] * i1;
        if (pass == 1) {
          y0 -= __bfloat162float(__float2bfloat16_rn(y0)); y1 -= __bfloat162float(__float2bfloat16_rn(y1));
          y2 -= __bfloat162float(__float2bfloat16_rn(y2)); y3 -= __bfloat162float(__float2bfloat16_rn(y3));
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qs + ra * (HD * 2) + ((j ^ (ra & 7)) << 4) + t4 * 4), "r"(pack_bf16x2(y0, y1)) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(qs + rb * (HD * 2) + ((j ^ (rb & 7)) << 4) + t4 * 4), "r"(pack_bf16x2(y2, y3)) : "memory");
      }
      __syncwarp();
      bf16* dst = pass == 0 ? out : out_lo;
      for (int i = lane; i < 16 * 8; i += 32) {
        const int tt = qb * 16 + (i >> 3), c = i & 7;
        if (tt < S) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(qs + tt * (HD * 2) + ((c ^ (tt & 7)) << 4)));
          *reinterpret_cast<uint4*>(dst + ((size_t)row * S + tt) * d + h * HD + c * 8) = v;
        }
      }
    }
  }
}

template <int NBLK, bool F16 = false>
static int launch_attn_prefill_mma(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, int B, int S, int H, int t_max, int cache_row_mult,
                                   cudaStream_t st, bf16* out_lo = nullptr) {
  const size_t smem = (size_t)4 * 3 * (16 * NBLK) * HD * 2 + 128;
  auto kern = attn_prefill_mma_kernel<NBLK, F16>;
  static std::atomic<bool> configured{false};
  if (!configured.load(std::memory_order_acquire) && smem > 48 * 1024) GIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  configured.store(true, std::memory_order_release);
  const int items = B * H;
  GIC_CHECK_CUDA(launch_kernel(kern, dim3(ceil_div(items, 4)), dim3(128), smem, st, qkv, kcache, vcache, out, items, S, H, t_max, cache_row_mult, out_lo));
  note_launch();
  return GIC_OK;
}

// causal prefill attention of the bf16x2 engine (fp16 q | k | v and cache, hi + lo output); prefixes longer than 64 tokens are not
// supported in this mode (the reference's configurations use 10 / 40 + a short task prompt)
int launch_attn_prefill_f16(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out_hi, bf16* out_lo, int B, int P, int H, int t_max, int cache_row_mult,
                            cudaStream_t st) {
  GIC_REQUIRE(P <= t_max && out_hi && out_lo, "attn_prefill_f16: bad arguments");
  GIC_REQUIRE(P <= 64, "bf16x2 engine: the prefix (image + task tokens) may hold at most 64 tokens, got %d", P);
  if (P <= 16) return launch_attn_prefill_mma<1, true>(qkv, kcache, vcache, out_hi, B, P, H, t_max, cache_row_mult, st, out_lo);
  if (P <= 32) return launch_attn_prefill_mma<2, true>(qkv, kcache, vcache, out_hi, B, P, H, t_max, cache_row_mult, st, out_lo);
  if (P <= 48) return launch_attn_prefill_mma<3, true>(qkv, kcache, vcache, out_hi, B, P, H, t_max, cache_row_mult, st, out_lo);
  return launch_attn_prefill_mma<4, true>(qkv, kcache, vcache, out_hi, B, P, H, t_max, cache_row_mult, st, out_lo);
}

template <typename T>
int launch_attn_prefill(const T* qkv, T* kcache, T* vcache, ActOut out, int B, int P, int H, int t_max, int cache_row_mult,
                        cudaStream_t st) {
  GIC_REQUIRE(P <= t_max, "attn_prefill: P %d > t_max %d", P, t_max);
  if (IsBf16<T>::value && out.hi && !out.lo && !out.f32 && kcache && P <= 64 && decode_bulk_enabled()) {
    const bf16* q = (const bf16*)qkv; bf16* kc = (bf16*)kcache; bf16* vc = (bf16*)vcache;
    if (P <= 16) return launch_attn_prefill_mma<1>(q, kc, vc, out.hi, B, P, H, t_max, cache_row_mult, st);
    if (P <= 32) return launch_attn_prefill_mma<2>(q, kc, vc, out.hi, B, P, H, t_max, cache_row_mult, st);
    if (P <= 48) return launch_attn_prefill_mma<3>(q, kc, vc, out.hi, B, P, H, t_max, cache_row_mult, st);
    return launch_attn_prefill_mma<4>(q, kc, vc, out.hi, B, P, H, t_max, cache_row_mult, st);
  }
  return launch_attn_seq<T, 64, true>(qkv, kcache, vcache, out, B, P, H, t_max, cache_row_mult, st);
}
template int launch_attn_prefill<float>(const float*, float*, float*, ActOut, int, int, int, int, int, cudaStream_t);
template int launch_attn_prefill<bf16>(const bf16*, bf16*, bf16*, ActOut, int, int, int, int, int, cudaStream_t);

template <typename T>
int launch_attn_encoder(const T* qkv, ActOut out, int B, int S, int H, int hd, cudaStream_t st) {
  switch (hd) {
    case 16: return launch_attn_seq<T, 16, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    case 32: return launch_attn_seq<T, 32, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    case 64: return launch_attn_seq<T, 64, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    case 96: return launch_attn_seq<T, 96, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    case 128: return launch_attn_seq<T, 128, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    case 160: return launch_attn_seq<T, 160, false>(qkv, nullptr, nullptr, out, B, S, H, 0, 1, st);
    default: set_error("attn_encoder: unsupported head_dim %d (16/32/64/96/128/160)", hd); return GIC_ERR_UNSUPPORTED;
  }
}
template int launch_attn_encoder<float>(const float*, ActOut, int, int, int, int, cudaStream_t);
template int launch_attn_encoder<bf16>(const bf16*, ActOut, int, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------------
// Beam reorder: dst[l][kv][r][h][0:ctx] = src[l][kv][beam_idx[r]][h][0:ctx].  One warp per (l, kv, r, h) segment,
// 128-bit loads/stores, only the live ctx_len positions are moved.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) kv_reorder_kernel(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ beam_idx,
                                                         int L2, int rows, int H, int ctx_len, int t_max) {
  constexpr int VEC = 16 / sizeof(T);
  const long seg = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nseg = (long)L2 * rows * H;
  if (seg >= nseg) return;
  const int lane = threadIdx.x & 31;
  const int h = (int)(seg % H);
  const int r = (int)((seg / H) % rows);
  const long l = seg / ((long)H * rows);
  const int sr = beam_idx[r];
  const T* s = src + (((size_t)l * rows + sr) * H + h) * t_max * HD;
  T* dd = dst + (((size_t)l * rows + r) * H + h) * t_max * HD;
  const int nvec = ctx_len * HD / VEC;
  for (int i = lane; i < nvec; i += 32) {
    Vec16<T> v;
    v.load(s + (size_t)i * VEC);
    v.store(dd + (size_t)i * VEC);
  }
}

template <typename T>
int launch_kv_reorder(const T* src, T* dst, const int* beam_idx, int L, int rows, int H, int ctx_len, int t_max, cudaStream_t st) {
  const long nseg = (long)L * 2 * rows * H;
  const int blocks = (int)((nseg + 7) / 8);
  kv_reorder_kernel<T><<<blocks, 256, 0, st>>>(src, dst, beam_idx, L * 2, rows, H, ctx_len, t_max);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}
template int launch_kv_reorder<float>(const float*, float*, const int*, int, int, int, int, int, cudaStream_t);
template int launch_kv_reorder<bf16>(const bf16*, bf16*, const int*, int, int, int, int, int, cudaStream_t);

}  // namespace gic
