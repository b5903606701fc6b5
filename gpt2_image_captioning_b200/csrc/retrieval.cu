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

}  // namespace gic
