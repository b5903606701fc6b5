// Retrieval for the RAT variant: exact inner-product top-k over the embedding database, the hit filter / caption-row
// selection and gather + aggregate + add.
// Replaces faiss IndexFlatIP.search + the Python loops of src/database/faiss_store.py:132-251 and
// RetrievalAggregator.forward (src/models.py:589-625); the reference does all of this on the CPU with a
// GPU->CPU->GPU hop per batch (src/models.py:677,695).
//
// top-k: the database is scanned once in chunks of TOPK_CHUNK rows: scores[B, chunk] = Q . DB_chunk^T (the fp32 GEMM of
// sgemm_fp32.cu -- exact fp32 products so indices match the fp32 reference), then one block per query keeps a running
// (score desc, index asc) top-k: per-thread sorted lists in registers, merged through shared memory.  The full [B, N]
// score matrix is never materialised.
#include <chrono>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace gic {

constexpr int TOPK_CHUNK = 32768;
constexpr int TOPK_MAXK = 32;

size_t topk_workspace_bytes(int B, int N, int D, int k) {
  (void)D; (void)k;
  const size_t chunk = (size_t)(N < TOPK_CHUNK ? N : TOPK_CHUNK);
  return align_up((size_t)B * chunk * sizeof(float), 256) + 256;
}

__device__ __forceinline__ bool cand_better(float v, long i, float bv, long bi) { return v > bv || (v == bv && i < bi); }

// the per-thread sorted lists of a block -> its k best (score desc, index asc) into scores / idx of row b
template <int KMAX>
__device__ __forceinline__ void topk_block_merge(const float (&lv)[KMAX], const long (&li)[KMAX], int k, float* cv, long* ci, float* rv, long* ri,
                                                 int* rslot, float* __restrict__ scores, int64_t* __restrict__ idx, int b) {
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < k) { cv[threadIdx.x * k + j] = lv[j]; ci[threadIdx.x * k + j] = li[j]; }
  __syncthreads();

  const int ncand = blockDim.x * k;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int out = 0; out < k; ++out) {
    float bv = -INFINITY; long bi = 0x7fffffffffffffffL; int bslot = -1;
    for (int s = threadIdx.x; s < ncand; s += blockDim.x)
      if (cand_better(cv[s], ci[s], bv, bi)) { bv = cv[s]; bi = ci[s]; bslot = s; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int os = __shfl_xor_sync(0xffffffffu, bslot, o);
      if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bslot = os; }
    }
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; rslot[warp] = bslot; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nw; ++w)
        if (cand_better(rv[w], ri[w], bv, bi)) { bv = rv[w]; bi = ri[w]; bslot = rslot[w]; }
      const bool valid = bslot >= 0 && bi != 0x7fffffffffffffffL;
      scores[(size_t)b * k + out] = valid ? bv : -INFINITY;
      idx[(size_t)b * k + out] = valid ? (int64_t)bi : -1;
      if (bslot >= 0) { cv[bslot] = -INFINITY; ci[bslot] = 0x7fffffffffffffffL; }
    }
    __syncthreads();
  }
}

template <int KMAX>
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ chunk_scores, int chunk_rows, long chunk_base, int k,
                                                         float* __restrict__ scores /*[B,k] running*/, int64_t* __restrict__ idx) {
  extern __shared__ unsigned char sm_raw[];
  float* cv = reinterpret_cast<float*>(sm_raw);                               // [nthreads * k]
  long* ci = reinterpret_cast<long*>(sm_raw + (size_t)blockDim.x * k * sizeof(float));  // [nthreads * k]
  __shared__ float rv[8];
  __shared__ long ri[8];
  __shared__ int rslot[8];
  const int b = blockIdx.x;
  const float* row = chunk_scores + (size_t)b * chunk_rows;

  float lv[KMAX];
  long li[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { lv[j] = -INFINITY; li[j] = 0x7fffffffffffffffL; }
  // the running result of earlier chunks joins as candidates (thread 0)
  if (threadIdx.x == 0) {
    for (int j = 0; j < k; ++j) {
      const int64_t gi = idx[(size_t)b * k + j];
      if (gi >= 0) { lv[j] = scores[(size_t)b * k + j]; li[j] = gi; }
    }
  }
  for (int c = threadIdx.x; c < chunk_rows; c += blockDim.x) {
    const float v = row[c];
    const long gi = chunk_base + c;
    float wv = -INFINITY; long wi = 0x7fffffffffffffffL;  // current worst kept entry (slot k-1), select chain keeps lv in registers
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j == k - 1) { wv = lv[j]; wi = li[j]; }
    if (cand_better(v, gi, wv, wi)) {
      // insertion into the sorted list (descending score, ascending index)
      float pv = v; long pi = gi;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < k && cand_better(pv, pi, lv[j], li[j])) {
          const float tv = lv[j]; const long ti = li[j];
          lv[j] = pv; li[j] = pi; pv = tv; pi = ti;
        }
      }
    }
  }
  topk_block_merge<KMAX>(lv, li, k, cv, ci, rv, ri, rslot, scores, idx, b);
}

__global__ void topk_init_kernel(float* scores, int64_t* idx, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { scores[i] = -INFINITY; idx[i] = -1; }
}

int launch_topk_ip(const float* q, const float* db, int B, int N, int D, int k, float* scores, int64_t* idx, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  GIC_REQUIRE(B > 0 && N > 0 && D > 0 && k > 0, "topk_ip: empty problem B=%d N=%d D=%d k=%d", B, N, D, k);
  GIC_REQUIRE(k <= TOPK_MAXK, "topk_ip: k=%d exceeds the supported maximum %d", k, TOPK_MAXK);
  GIC_REQUIRE(D % 4 == 0, "topk_ip: D (%d) must be a multiple of 4", D);
  GIC_REQUIRE(ws_bytes >= topk_workspace_bytes(B, N, D, k), "topk_ip: workspace too small");
  float* chunk_scores = reinterpret_cast<float*>(align_up((size_t)ws, 256));
  const size_t n = (size_t)B * k;
  topk_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, idx, n);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  const int nthreads = k <= 8 ? 256 : (k <= 16 ? 128 : 64);  // candidate lists must fit 48 KB of shared memory
  const size_t smem = (size_t)nthreads * k * (sizeof(float) + sizeof(long));
  for (long base = 0; base < N; base += TOPK_CHUNK) {
    const int rows = (int)((N - base) < TOPK_CHUNK ? (N - base) : TOPK_CHUNK);
    GIC_TRY(launch_sgemm_nt(q, D, db + (size_t)base * D, nullptr, chunk_scores, rows, B, rows, D, EPI_NONE, st));
    if (k <= 8) topk_merge_kernel<8><<<B, nthreads, smem, st>>>(chunk_scores, rows, base, k, scores, idx);
    else if (k <= 16) topk_merge_kernel<16><<<B, nthreads, smem, st>>>(chunk_scores, rows, base, k, scores, idx);
    else topk_merge_kernel<32><<<B, nthreads, smem, st>>>(chunk_scores, rows, base, k, scores, idx);
    GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  }
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core top-k with the exact path's results.  The fp32 scan above runs at ~27 TFLOP/s on the CUDA cores (23 ms for
// 1024 queries over the 591 753-row caption matrix, profiles/r1ak_configs.jsonl).  Here the scan is a bf16x2 tcgen05 GEMM
// (operands split hi + lo, three MMAs per product: error <= ~2.3e-5 |q||d|), which only has to find CANDIDATES:
//   1. ONE GEMM launch over the whole database whose epilogue never writes a score (EPI_TOPK, gemm_tcgen05.cu): CTAs are
//      row-tile-stationary, every epilogue thread owns one query row and keeps, over all the column tiles it walks, its 8 best
//      (score, row id) and the largest score it DROPPED; a query ends with 2 x (CTAs per row tile) such streams.  The 32 best of the
//      survivors become the candidates and U = the largest score dropped anywhere.  (Round 1 materialised [B, 32 768] fp32 score
//      chunks and re-read them with a scan kernel: 4.8 GB of traffic for configs[4], 3.7 ms; fused: the GEMM's own time.)
//   2. every candidate is re-scored exactly: one thread per candidate, fp32 FMA chain over the dims in ascending order --
//      the summation order of sgemm_nt_kernel, so scores AND ranks are bit-identical to the exact path;
//   3. certificate: U bounds every non-candidate's approximate score, so its exact score is at most U + eps,
//      eps = 6e-5 |q| max|d|.  If the k-th exact score is above that, no outsider can enter (or tie into) the top k.
//      Otherwise the row is flagged and ONE block rescans the whole database for it exactly (no host round trip; with
//      spacings of ~1e-3 between the leading scores of a 10^5..10^6-row database this practically never runs).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TOPK_TC_CAND = 32;
constexpr int TOPK_TC_MAXK = 16;
constexpr int TOPK_TC_THREADS = 256;   // select kernel threads per query
constexpr int TOPK_TC_KEEP = GEMM_TOPK_KEEP_PUBLIC;  // survivors per (query, stream): the fused epilogue's running list
constexpr int TOPK_TC_MAX_STREAMS = 2 * 160;  // 2 per CTA of a row tile, at most one row tile and every SM

size_t topk_tc_workspace_bytes(int B, int N, int D, int k) {
  (void)k; (void)N;
  const size_t surv = (size_t)B * TOPK_TC_MAX_STREAMS * TOPK_TC_KEEP;
  return 2 * align_up((size_t)B * D * sizeof(bf16), 256) + align_up((size_t)B * TOPK_TC_CAND * sizeof(float), 256) +
         align_up((size_t)B * TOPK_TC_CAND * sizeof(int64_t), 256) + align_up(((size_t)B + 1) * sizeof(int), 256) +  // flags [B] + the number of flagged queries
         align_up(surv * sizeof(float), 256) + align_up(surv * sizeof(int), 256) + align_up((size_t)B * TOPK_TC_MAX_STREAMS * sizeof(float), 256) +
         align_up((size_t)B * sizeof(float), 256) + 512;  // scan state + bounds
}

// exact fp32 inner product in the summation order of the exact path (ascending k, one FMA chain from 0)
__device__ __forceinline__ float dot_exact(const float* __restrict__ q_s, const float* __restrict__ row, int D) {
  float acc = 0.f;
  for (int k = 0; k < D; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + k);
    acc = fmaf(q_s[k], v.x, acc);
    acc = fmaf(q_s[k + 1], v.y, acc);
    acc = fmaf(q_s[k + 2], v.z, acc);
    acc = fmaf(q_s[k + 3], v.w, acc);
  }
  return acc;
}

// the TOPK_TC_CAND best of a query's `streams` x TOPK_TC_KEEP survivors -> candidate list; bound[b] = the largest approximate score that is
// NOT a candidate (dropped by a stream, or a survivor that did not make the list).  A query would be wrong only if one stream held more than
// TOPK_TC_KEEP of the rows that matter and dropped one -- which the certificate below catches (the dropped score raises U).
__global__ void __launch_bounds__(TOPK_TC_THREADS) topk_tc_select_kernel(const float* __restrict__ st_v, const int* __restrict__ st_i,
                                                                         const float* __restrict__ st_u, int streams, float* __restrict__ cand_score,
                                                                         int64_t* __restrict__ cand_idx, float* __restrict__ bound) {
  extern __shared__ unsigned char sel_raw[];
  const int S = streams * TOPK_TC_KEEP;
  float* sv = reinterpret_cast<float*>(sel_raw);  // [S]
  int* si = reinterpret_cast<int*>(sv + S);        // [S]
  __shared__ float rv[8];
  __shared__ int rs[8];
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < S; i += TOPK_TC_THREADS) { sv[i] = st_v[(size_t)b * S + i]; si[i] = st_i[(size_t)b * S + i]; }
  float u = -INFINITY;
  for (int i = t; i < streams; i += TOPK_TC_THREADS) u = fmaxf(u, st_u[(size_t)b * streams + i]);
  __syncthreads();
  for (int out = 0; out < TOPK_TC_CAND; ++out) {
    // block argmax over the survivors still in the pool (ties: the lower row id, so that the candidates do not depend on the stream layout)
    float bv = -INFINITY; int bs = -1, bid = 0x7fffffff;
    for (int i = t; i < S; i += TOPK_TC_THREADS) {
      const float v = sv[i]; const int id = si[i];
      if (id >= 0 && (v > bv || (v == bv && id < bid))) { bv = v; bs = i; bid = id; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oid = __shfl_xor_sync(0xffffffffu, bid, o);
      if (os >= 0 && (bs < 0 || ov > bv || (ov == bv && oid < bid))) { bv = ov; bs = os; bid = oid; }
    }
    if (lane == 0) { rv[warp] = bv; rs[warp] = bs; }
    __syncthreads();
    if (t == 0) {
      int best = -1; float best_v = -INFINITY; int best_id = 0x7fffffff;
      for (int w = 0; w < TOPK_TC_THREADS / 32; ++w) {
        const int sidx = rs[w];
        if (sidx < 0) continue;
        const float v = rv[w]; const int id = si[sidx];
        if (best < 0 || v > best_v || (v == best_v && id < best_id)) { best = sidx; best_v = v; best_id = id; }
      }
      cand_score[(size_t)b * TOPK_TC_CAND + out] = best >= 0 ? best_v : -INFINITY;
      cand_idx[(size_t)b * TOPK_TC_CAND + out] = best >= 0 ? (int64_t)best_id : -1;
      if (best >= 0) { sv[best] = -INFINITY; si[best] = -1; }
    }
    __syncthreads();
  }
  for (int i = t; i < S; i += TOPK_TC_THREADS)
    if (si[i] >= 0) u = fmaxf(u, sv[i]);  // survivors that did not make the list count as dropped
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
  if (lane == 0) rv[warp] = u;
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < TOPK_TC_THREADS / 32; ++w) u = fmaxf(u, rv[w]);
    bound[b] = u;
  }
}

// one block (TOPK_TC_CAND threads) per query: exact scores of its candidates, exact (score desc, index asc) top k, certificate
__global__ void __launch_bounds__(TOPK_TC_CAND) topk_rescore_kernel(const float* __restrict__ q, const float* __restrict__ db, int D, int k,
                                                                    const int64_t* __restrict__ cand_idx, const float* __restrict__ bound,
                                                                    float db_norm_max, float* __restrict__ scores, int64_t* __restrict__ idx,
                                                                    int* __restrict__ flags) {
  extern __shared__ float q_s[];  // [D]
  __shared__ float ev[TOPK_TC_CAND];
  __shared__ long ei[TOPK_TC_CAND];
  __shared__ float s_kth;
  const int b = blockIdx.x, j = threadIdx.x;
  float qq = 0.f;
  for (int c = j; c < D; c += TOPK_TC_CAND) { const float v = q[(size_t)b * D + c]; q_s[c] = v; qq += v * v; }
  qq = warp_sum(qq);  // the block is one warp
  __syncwarp();
  const int64_t gi = cand_idx[(size_t)b * TOPK_TC_CAND + j];
  const float e = gi >= 0 ? dot_exact(q_s, db + (size_t)gi * D, D) : -INFINITY;
  ev[j] = e;
  ei[j] = gi >= 0 ? (long)gi : 0x7fffffffffffffffL;
  if (j == 0) s_kth = -INFINITY;
  __syncwarp();
  int rank = 0;
  for (int i = 0; i < TOPK_TC_CAND; ++i)  // (empty slots are all alike: ordered by slot)
    rank += (cand_better(ev[i], ei[i], e, ei[j]) || (ev[i] == e && ei[i] == ei[j] && i < j)) ? 1 : 0;
  if (rank < k) {
    scores[(size_t)b * k + rank] = gi >= 0 ? e : -INFINITY;
    idx[(size_t)b * k + rank] = gi;
    if (rank == k - 1) s_kth = e;
  }
  __syncwarp();
  if (j == 0) {
    // bound[b] = the largest approximate score of any row outside the candidate list (-inf when the list holds every row)
    const float u = bound[b];
    const float eps = 6e-5f * sqrtf(qq) * db_norm_max;
    flags[b] = (u > -INFINITY && !(s_kth > u + eps)) ? 1 : 0;
  }
}

// flagged rows only: exact scan of the whole database by one block (the exact path's scores and tie rule)
template <int KMAX>
__global__ void __launch_bounds__(256) topk_exact_row_kernel(const float* __restrict__ q, const float* __restrict__ db, long N, int D, int k,
                                                             const int* __restrict__ flags, float* __restrict__ scores, int64_t* __restrict__ idx) {
  extern __shared__ unsigned char sm_raw[];
  const int b = blockIdx.x;
  if (!flags[b]) return;
  if (threadIdx.x == 0) atomicAdd(const_cast<int*>(flags) + gridDim.x, 1);  // diagnostics: flags[B] counts the rescanned queries
  float* q_s = reinterpret_cast<float*>(sm_raw);                                        // [D]
  float* cv = q_s + D;                                                                   // [nthreads * k]
  long* ci = reinterpret_cast<long*>(sm_raw + ((((size_t)D + (size_t)blockDim.x * k) * sizeof(float) + 7) & ~(size_t)7));  // [nthreads * k]
  __shared__ float rv[8];
  __shared__ long ri[8];
  __shared__ int rslot[8];
  for (int c = threadIdx.x; c < D; c += blockDim.x) q_s[c] = q[(size_t)b * D + c];
  __syncthreads();
  float lv[KMAX];
  long li[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { lv[j] = -INFINITY; li[j] = 0x7fffffffffffffffL; }
  for (long c = threadIdx.x; c < N; c += blockDim.x) {
    const float v = dot_exact(q_s, db + (size_t)c * D, D);
    float wv = -INFINITY; long wi = 0x7fffffffffffffffL;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j == k - 1) { wv = lv[j]; wi = li[j]; }
    if (cand_better(v, c, wv, wi)) {
      float pv = v; long pi = c;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < k && cand_better(pv, pi, lv[j], li[j])) {
          const float tv = lv[j]; const long ti = li[j];
          lv[j] = pv; li[j] = pi; pv = tv; pi = ti;
        }
      }
    }
  }
  topk_block_merge<KMAX>(lv, li, k, cv, ci, rv, ri, rslot, scores, idx, b);
}

bool topk_tc_supported(int D, int k) { return D % 64 == 0 && D <= 2048 && k <= TOPK_TC_MAXK; }

int launch_topk_ip_tc(const float* q, const float* db, const bf16* db_hi, const bf16* db_lo, float db_norm_max, int B, int N, int D, int k,
                      float* scores, int64_t* idx, void* ws, size_t ws_bytes, cudaStream_t st) {
  GIC_REQUIRE(B > 0 && N > 0 && D > 0 && k > 0, "topk_ip_tc: empty problem B=%d N=%d D=%d k=%d", B, N, D, k);
  GIC_REQUIRE((long)N < 0x7fffffffL, "topk_ip_tc: N too large");
  GIC_REQUIRE(topk_tc_supported(D, k), "topk_ip_tc: needs D %% 64 == 0, D <= 2048 and k <= %d (D=%d k=%d); use the exact path", TOPK_TC_MAXK, D, k);
  GIC_REQUIRE(ws_bytes >= topk_tc_workspace_bytes(B, N, D, k), "topk_ip_tc: workspace too small");
  GIC_TRY(tma_init());
  GIC_TRY(gemm_bf16_configure());
  unsigned char* w8 = reinterpret_cast<unsigned char*>(align_up((size_t)ws, 256));
  bf16* q_hi = reinterpret_cast<bf16*>(w8); w8 += align_up((size_t)B * D * sizeof(bf16), 256);
  bf16* q_lo = reinterpret_cast<bf16*>(w8); w8 += align_up((size_t)B * D * sizeof(bf16), 256);
  float* cand_score = reinterpret_cast<float*>(w8); w8 += align_up((size_t)B * TOPK_TC_CAND * sizeof(float), 256);
  int64_t* cand_idx = reinterpret_cast<int64_t*>(w8); w8 += align_up((size_t)B * TOPK_TC_CAND * sizeof(int64_t), 256);
  int* flags = reinterpret_cast<int*>(w8); w8 += align_up(((size_t)B + 1) * sizeof(int), 256);
  const size_t surv_max = (size_t)B * TOPK_TC_MAX_STREAMS * TOPK_TC_KEEP;
  float* st_v = reinterpret_cast<float*>(w8); w8 += align_up(surv_max * sizeof(float), 256);
  int* st_i = reinterpret_cast<int*>(w8); w8 += align_up(surv_max * sizeof(int), 256);
  float* st_u = reinterpret_cast<float*>(w8); w8 += align_up((size_t)B * TOPK_TC_MAX_STREAMS * sizeof(float), 256);
  float* bound = reinterpret_cast<float*>(w8);

  ActOut qo; qo.hi = q_hi; qo.lo = q_lo;
  GIC_TRY(launch_convert(q, qo, (size_t)B * D, st));
  // one bf16x2 GEMM over the whole database, scores consumed in the epilogue
  GemmBf16Args g;
  g.no_pdl = 1;  // plain stream order: its neighbours are ordinary <<<>>> launches on the caller's stream
  int bn = 128, pair = 0;
  gemm_bf16_pick(B, N, D, 1, 1, &bn, &pair);
  if (!pair) bn = 128;  // (single CTAs: the fused epilogue exists for the 128-column bf16x2 tile)
  GIC_TRY(make_tma_2d_bf16(&g.a_hi, q_hi, B, D, D, 128));
  GIC_TRY(make_tma_2d_bf16(&g.a_lo, q_lo, B, D, D, 128));
  GIC_TRY(make_tma_2d_bf16(&g.w_hi, db_hi, N, D, D, pair ? bn / 2 : bn));
  GIC_TRY(make_tma_2d_bf16(&g.w_lo, db_lo, N, D, D, pair ? bn / 2 : bn));
  const int streams = gemm_topk_streams(B, N, bn, pair);
  GIC_REQUIRE(streams <= TOPK_TC_MAX_STREAMS, "topk_ip_tc: %d candidate streams exceed the workspace layout", streams);
  g.M = B; g.N = N; g.K = D; g.block_n = bn; g.pair = pair; g.split = 1; g.epilogue = EPI_NONE; g.bias = nullptr;
  g.topk_v = st_v; g.topk_i = st_i; g.topk_u = st_u; g.topk_streams = streams;
  GIC_TRY(launch_gemm_bf16(g, st));
  const size_t sel_smem = (size_t)streams * TOPK_TC_KEEP * (sizeof(float) + sizeof(int));
  topk_tc_select_kernel<<<B, TOPK_TC_THREADS, sel_smem, st>>>(st_v, st_i, st_u, streams, cand_score, cand_idx, bound);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  GIC_CHECK_CUDA(cudaMemsetAsync(flags + B, 0, sizeof(int), st));
  topk_rescore_kernel<<<B, TOPK_TC_CAND, (size_t)D * sizeof(float), st>>>(q, db, D, k, cand_idx, bound, db_norm_max, scores, idx, flags);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  const int fthreads = k <= 8 ? 256 : 128;  // candidate lists + the query must fit 48 KB of shared memory
  const size_t fsmem = align_up(((size_t)D + fthreads * (size_t)k) * sizeof(float), 8) + fthreads * (size_t)k * sizeof(long);
  GIC_REQUIRE(fsmem <= 48 * 1024, "topk_ip_tc: fix-up kernel shared memory %zu", fsmem);
  topk_exact_row_kernel<TOPK_TC_MAXK><<<B, fthreads, fsmem, st>>>(q, db, (long)N, D, k, flags, scores, idx);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// faiss_store.py:160-183 (hit filter) + :208-229 (caption-row selection), on integer ids
__global__ void select_caption_rows_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int B, int k_searched,
                                           const int64_t* __restrict__ cap_row_start, const int64_t* __restrict__ cap_row_ids,
                                           int top_i, int top_k, int64_t* __restrict__ rows_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t* out = rows_out + (size_t)b * top_k;
  int n_rows = 0, n_hits = 0;
  for (int j = 0; j < k_searched && n_hits < top_i; ++j) {
    const int64_t img = idx[(size_t)b * k_searched + j];
    if (img < 0) continue;                                   // idx == -1
    if (scores[(size_t)b * k_searched + j] > 0.9999f) continue;  // near-perfect match = the query image itself
    ++n_hits;
    // the reference stops extending once it holds >= top_k rows (:217-219) and keeps the first top_k (:229)
    for (int64_t r = cap_row_start[img]; r < cap_row_start[img + 1] && n_rows < top_k; ++r) out[n_rows++] = cap_row_ids ? cap_row_ids[r] : r;
  }
  for (int j = n_rows; j < top_k; ++j) out[j] = -1;
}

int launch_select_caption_rows(const float* scores, const int64_t* idx, int B, int k_searched, const int64_t* cap_row_start,
                               const int64_t* cap_row_ids, int top_i, int top_k, int64_t* rows_out, cudaStream_t st) {
  GIC_REQUIRE(B > 0 && top_k > 0 && top_i > 0, "select_caption_rows: bad sizes");
  select_caption_rows_kernel<<<ceil_div(B, 128), 128, 0, st>>>(scores, idx, B, k_searched, cap_row_start, cap_row_ids, top_i, top_k, rows_out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// out[b] = q[b] + agg_k(cap_db[rows[b,k]]) -- zero rows for -1 (they count in the mean / max, src/models.py:591-595)
__global__ void __launch_bounds__(256) gather_aggregate_add_kernel(const float* __restrict__ q, const float* __restrict__ cap_db,
                                                                   const int64_t* __restrict__ rows, int top_k, int D, int aggregation,
                                                                   float* __restrict__ out) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int64_t* r = rows + (size_t)b * top_k;
  constexpr int MAXC = 8;  // D <= 2048
  float acc[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) acc[i] = (aggregation == GIC_AGG_MAX) ? -INFINITY : 0.f;
  for (int j = 0; j < top_k; ++j) {
    const int64_t ri = r[j];
    float v[MAXC];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = threadIdx.x + i * 256;
      v[i] = (ri >= 0 && c < D) ? cap_db[(size_t)ri * D + c] : 0.f;
      ss += v[i] * v[i];
    }
    float scale = 1.f;
    if (aggregation == GIC_AGG_SUM_NORM) scale = 1.f / fmaxf(sqrtf(block_sum_256(ss, red)), 1e-12f);  // F.normalize eps
#pragma unroll
    for (int i = 0; i < MAXC; ++i) acc[i] = (aggregation == GIC_AGG_MAX) ? fmaxf(acc[i], v[i]) : acc[i] + v[i] * scale;
  }
  float post = 1.f;
  if (aggregation == GIC_AGG_MEAN) post = 1.f / (float)top_k;
  if (aggregation == GIC_AGG_SUM_NORM) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) ss += acc[i] * acc[i];
    post = 1.f / fmaxf(sqrtf(block_sum_256(ss, red)), 1e-12f);
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < D) {
      const float a = (aggregation == GIC_AGG_MEAN) ? acc[i] / (float)top_k : acc[i] * post;
      out[(size_t)b * D + c] = q[(size_t)b * D + c] + a;
    }
  }
}

int launch_gather_aggregate_add(const float* q, const float* cap_db, const int64_t* rows, int B, int top_k, int D, int aggregation,
                                float* out, cudaStream_t st) {
  GIC_REQUIRE(D <= 2048 && D > 0 && top_k > 0, "gather_aggregate_add: unsupported D=%d top_k=%d", D, top_k);
  GIC_REQUIRE(aggregation >= GIC_AGG_MEAN && aggregation <= GIC_AGG_SUM_NORM, "gather_aggregate_add: unknown aggregation %d", aggregation);
  gather_aggregate_add_kernel<<<B, 256, 0, st>>>(q, cap_db, rows, top_k, D, aggregation, out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// get_caption_embeddings' reconstruct loop + zero padding (faiss_store.py:229-251) as one gather: out[b, j, :] = cap_db[rows[b, j]] or 0
__global__ void __launch_bounds__(128) gather_caption_rows_kernel(const float* __restrict__ cap_db, const int64_t* __restrict__ rows, int D,
                                                                  float* __restrict__ out) {
  const int64_t ri = rows[blockIdx.x];
  float* dst = out + (size_t)blockIdx.x * D;
  const float* src = cap_db + (size_t)(ri < 0 ? 0 : ri) * D;
  if ((D & 3) == 0) {
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
      *reinterpret_cast<float4*>(dst + c) = ri >= 0 ? *reinterpret_cast<const float4*>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int c = threadIdx.x; c < D; c += blockDim.x) dst[c] = ri >= 0 ? src[c] : 0.f;
  }
}

int launch_gather_caption_rows(const float* cap_db, const int64_t* rows, int n_rows, int D, float* out, cudaStream_t st) {
  GIC_REQUIRE(D > 0 && n_rows >= 0, "gather_caption_rows: bad sizes");
  if (n_rows == 0) return GIC_OK;
  gather_caption_rows_kernel<<<n_rows, 128, 0, st>>>(cap_db, rows, D, out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// RetrievalAggregator "attention" (src/models.py:606-616): score_j = w . r_j + b over the top_k gathered rows (zero rows for -1 padding,
// whose score is b), softmax over j, out = q + sum_j weight_j r_j
__global__ void __launch_bounds__(256) gather_attention_add_kernel(const float* __restrict__ q, const float* __restrict__ cap_db,
                                                                   const int64_t* __restrict__ rows, int top_k, int D, const float* __restrict__ attn_w,
                                                                   const float* __restrict__ attn_b, float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float sc[64];
  const int b = blockIdx.x;
  const int64_t* r = rows + (size_t)b * top_k;
  constexpr int MAXC = 8;  // D <= 2048
  float w[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) { const int c = threadIdx.x + i * 256; w[i] = c < D ? attn_w[c] : 0.f; }
  const float bias = attn_b[0];
  for (int j = 0; j < top_k; ++j) {
    const int64_t ri = r[j];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = threadIdx.x + i * 256;
      dot += (ri >= 0 && c < D) ? cap_db[(size_t)ri * D + c] * w[i] : 0.f;
    }
    const float s = block_sum_256(dot, red) + bias;
    if (threadIdx.x == 0) sc[j] = s;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = 0; j < top_k; ++j) m = fmaxf(m, sc[j]);
  float den = 0.f;
  for (int j = 0; j < top_k; ++j) den += expf(sc[j] - m);
  float acc[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) acc[i] = 0.f;
  for (int j = 0; j < top_k; ++j) {
    const int64_t ri = r[j];
    const float wt = expf(sc[j] - m) / den;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = threadIdx.x + i * 256;
      if (ri >= 0 && c < D) acc[i] = fmaf(wt, cap_db[(size_t)ri * D + c], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < D) out[(size_t)b * D + c] = q[(size_t)b * D + c] + acc[i];
  }
}

int launch_gather_attention_add(const float* q, const float* cap_db, const int64_t* rows, int B, int top_k, int D, const float* attn_w,
                                const float* attn_b, float* out, cudaStream_t st) {
  GIC_REQUIRE(D <= 2048 && D > 0 && top_k > 0 && top_k <= 64, "gather_attention_add: unsupported D=%d top_k=%d (top_k <= 64)", D, top_k);
  gather_attention_add_kernel<<<B, 256, 0, st>>>(q, cap_db, rows, top_k, D, attn_w, attn_b, out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

}  // namespace gic
