// Beam search on the device.  The reference has no beam search; semantics follow HF
// GenerationMixin._beam_search (HF:generation/utils.py:3076-3385) with do_sample=False, early_stopping=False,
// num_return_sequences=1, as invoked by the oracle:
//   per step: log_softmax(fp32 logits) + running beam score -> top 2*beams over beams x V (:2945-2997) ->
//   EOS / max-length candidates leave the running set (:2999-3019) and the top `beams` of them may enter the finished
//   pool with score / len^length_penalty (:3021-3074) -> KV cache reordered by the surviving beams' parents ->
//   early-stop heuristic (:2876-2921).
// Two kernels per step: beam_topk (one block per image: log-sum-exp of every live beam row, then the 2*beams best
// continuations; ties -> lowest flat index) and beam_update (one thread per image: the small-array bookkeeping),
// plus beam_embed (next input rows) and the KV gather of attention.cu.
#include "kernels.cuh"

namespace gic {

__device__ __forceinline__ bool bcand_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// logits: [B * rows_per_image, V] fp32 (rows_per_image = beams; = 1 at step 0 where only beam 0 is live).
// Three launches, one block per LIVE ROW for the two passes over the logits (the first version ran one block per image with a
// single dependent load in flight per thread: 4.5 ms per step for B = 1024 x 5 beams, 20x the time the 2 GB of reads need):
//   beam_lse_kernel      log-sum-exp of a row (eight independent loads in flight per thread)
//   beam_rowtopk_kernel  the row's K best continuations by (logit - lse) + running score, ties -> lowest flat index j V + c
//   beam_merge_kernel    per image: the K best of its n_live x K row candidates (an image's top K is inside its rows' top Ks)
constexpr int BEAM_UNROLL = 8;

__global__ void __launch_bounds__(256) beam_lse_kernel(const float* __restrict__ logits, int rows_per_image, int n_live, int V,
                                                       float* __restrict__ lse /* [B * rows_per_image] */) {
  __shared__ float red_m[8], red_s[8];
  const int b = blockIdx.x / n_live, j = blockIdx.x % n_live;
  const float* row = logits + ((size_t)b * rows_per_image + j) * V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = -INFINITY, s = 0.f;
  for (int c0 = threadIdx.x; c0 < V; c0 += blockDim.x * BEAM_UNROLL) {
    float x[BEAM_UNROLL];
#pragma unroll
    for (int u = 0; u < BEAM_UNROLL; ++u) {
      const int c = c0 + u * blockDim.x;
      x[u] = c < V ? row[c] : -INFINITY;
    }
    float bm = x[0];
#pragma unroll
    for (int u = 1; u < BEAM_UNROLL; ++u) bm = fmaxf(bm, x[u]);
    const float nm = fmaxf(m, bm);  // finite: x[0] is a real logit
    float add = 0.f;
#pragma unroll
    for (int u = 0; u < BEAM_UNROLL; ++u) add += expf(x[u] - nm);  // exp(-inf) = 0 for the padding slots
    s = (m == -INFINITY ? 0.f : s * expf(m - nm)) + add;
    m = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (m == -INFINITY ? 0.f : s * expf(m - nm)) + (om == -INFINITY ? 0.f : os * expf(om - nm));
    m = nm;
  }
  if (lane == 0) { red_m[warp] = m; red_s[warp] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = red_m[0], S = red_s[0];
    for (int w = 1; w < nw; ++w) {
      const float nm = fmaxf(M, red_m[w]);
      S = (M == -INFINITY ? 0.f : S * expf(M - nm)) + (red_m[w] == -INFINITY ? 0.f : red_s[w] * expf(red_m[w] - nm));
      M = nm;
    }
    lse[(size_t)b * rows_per_image + j] = M + logf(S);
  }
}

// the block's per-thread sorted lists -> its K best (value desc, index asc) into out_v / out_i
template <int K>
__device__ __forceinline__ void beam_block_merge(const float (&lv)[K], const int (&li)[K], float* cv, int* ci, float* rv, int* ri, int* rslot,
                                                 float* __restrict__ out_v, int* __restrict__ out_i) {
#pragma unroll
  for (int t = 0; t < K; ++t) { cv[threadIdx.x * K + t] = lv[t]; ci[threadIdx.x * K + t] = li[t]; }
  __syncthreads();
  const int ncand = blockDim.x * K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int out = 0; out < K; ++out) {
    float bv = -INFINITY; int bi = 0x7fffffff, bslot = -1;
    for (int s = threadIdx.x; s < ncand; s += blockDim.x)
      if (bcand_better(cv[s], ci[s], bv, bi)) { bv = cv[s]; bi = ci[s]; bslot = s; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), os = __shfl_xor_sync(0xffffffffu, bslot, o);
      if (bcand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bslot = os; }
    }
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; rslot[warp] = bslot; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nw; ++w)
        if (bcand_better(rv[w], ri[w], bv, bi)) { bv = rv[w]; bi = ri[w]; bslot = rslot[w]; }
      out_v[out] = bv;
      out_i[out] = bi;
      if (bslot >= 0) { cv[bslot] = -INFINITY; ci[bslot] = 0x7fffffff; }
    }
    __syncthreads();
  }
}

template <int K>
__global__ void __launch_bounds__(256) beam_rowtopk_kernel(const float* __restrict__ logits, int rows_per_image, int n_live,
                                                           const float* __restrict__ run_score, const float* __restrict__ lse, int beams, int V,
                                                           float* __restrict__ row_val, int* __restrict__ row_idx /* [B * rows_per_image, K] */) {
  extern __shared__ unsigned char sm_raw[];
  float* cv = reinterpret_cast<float*>(sm_raw);                                  // [256 * K]
  int* ci = reinterpret_cast<int*>(sm_raw + (size_t)blockDim.x * K * sizeof(float));  // [256 * K]
  __shared__ float rv[8];
  __shared__ int ri[8], rslot[8];
  const int b = blockIdx.x / n_live, j = blockIdx.x % n_live;
  const size_t r = (size_t)b * rows_per_image + j;
  const float* row = logits + r * V;
  const float base = run_score[(size_t)b * beams + j], l = lse[r];
  float lv[K];
  int li[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { lv[t] = -INFINITY; li[t] = 0x7fffffff; }
  for (int c0 = threadIdx.x; c0 < V; c0 += blockDim.x * BEAM_UNROLL) {
    float x[BEAM_UNROLL];
#pragma unroll
    for (int u = 0; u < BEAM_UNROLL; ++u) {
      const int c = c0 + u * blockDim.x;
      x[u] = c < V ? row[c] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < BEAM_UNROLL; ++u) {
      const int c = c0 + u * blockDim.x;
      const float v = (x[u] - l) + base;  // log_softmax first, then + running score (HF :3252-3256,3283)
      const int gi = j * V + c;
      if (c < V && bcand_better(v, gi, lv[K - 1], li[K - 1])) {
        float pv = v; int pi = gi;
#pragma unroll
        for (int t = 0; t < K; ++t)
          if (bcand_better(pv, pi, lv[t], li[t])) {
            const float tv = lv[t]; const int ti = li[t];
            lv[t] = pv; li[t] = pi; pv = tv; pi = ti;
          }
      }
    }
  }
  beam_block_merge<K>(lv, li, cv, ci, rv, ri, rslot, row_val + r * K, row_idx + r * K);
}

// one warp per image: the K best of its n_live x K row candidates
__global__ void __launch_bounds__(32) beam_merge_kernel(const float* __restrict__ row_val, const int* __restrict__ row_idx, int rows_per_image,
                                                        int n_live, int K, float* __restrict__ cand_score, int* __restrict__ cand_idx) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = n_live * K;  // <= 8 * 16 = 128 candidates: up to 4 per lane
  float v[4];
  int ix[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int s = lane + 32 * u;
    const bool ok = s < n;
    const size_t src = ((size_t)b * rows_per_image + (ok ? s / K : 0)) * K + (ok ? s % K : 0);
    v[u] = ok ? row_val[src] : -INFINITY;
    ix[u] = ok ? row_idx[src] : 0x7fffffff;
  }
  for (int out = 0; out < K; ++out) {
    float bv = v[0]; int bi = ix[0], bu = 0;
#pragma unroll
    for (int u = 1; u < 4; ++u)
      if (bcand_better(v[u], ix[u], bv, bi)) { bv = v[u]; bi = ix[u]; bu = u; }
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ol = __shfl_xor_sync(0xffffffffu, bl, o), ou = __shfl_xor_sync(0xffffffffu, bu, o);
      if (bcand_better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; bu = ou; }
    }
    if (lane == 0) { cand_score[(size_t)b * K + out] = bv; cand_idx[(size_t)b * K + out] = bi; }
    if (lane == bl) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u == bu) { v[u] = -INFINITY; ix[u] = 0x7fffffff; }
    }
  }
}

template <int K>
static int launch_beam_rowtopk(const float* logits, int B, int rows_per_image, int n_live, const float* run_score, const float* lse, int beams,
                               int V, float* row_val, int* row_idx, cudaStream_t st) {
  const size_t smem = (size_t)256 * K * (sizeof(float) + sizeof(int));
  beam_rowtopk_kernel<K><<<B * n_live, 256, smem, st>>>(logits, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// ---- the same step without a logits matrix: the LM-head GEMM's EPI_BEAM epilogue (gemm_tcgen05.cu) leaves, per row and stream, the stream's
// GEMM_BEAM_KEEP best (logit, column) pairs and its running (maximum, sum of exp(logit - maximum)).  One warp per live row combines the
// streams' log-sum-exp partials in stream order, turns the kept logits into (logit - lse) + running score exactly as beam_rowtopk_kernel
// does, and selects the row's K best (value desc, flat index asc).  A row's top K by value lies inside the union of its streams' top 16 by
// logit (K <= 16; value is a non-decreasing function of the logit). ----
__global__ void __launch_bounds__(32) beam_streams_kernel(const float* __restrict__ tk_v, const int* __restrict__ tk_i, const float* __restrict__ bm,
                                                          const float* __restrict__ bs, int streams, int rows_per_image, int n_live,
                                                          const float* __restrict__ run_score, int beams, int V, int K, float* __restrict__ lse,
                                                          float* __restrict__ row_val, int* __restrict__ row_idx) {
  extern __shared__ unsigned char sm_raw[];
  const int n = streams * GEMM_BEAM_KEEP;
  float* cv = reinterpret_cast<float*>(sm_raw);             // [n]
  int* ci = reinterpret_cast<int*>(sm_raw + (size_t)n * 4);  // [n]
  const int b = blockIdx.x / n_live, j = blockIdx.x % n_live, lane = threadIdx.x;
  const size_t r = (size_t)b * rows_per_image + j;
  // log-sum-exp: every lane runs the same short loop (streams <= a few dozen): no reduction order to worry about
  float M = -INFINITY;
  for (int s = 0; s < streams; ++s) M = fmaxf(M, bm[r * streams + s]);
  float S = 0.f;
  for (int s = 0; s < streams; ++s) {
    const float m = bm[r * streams + s];
    if (m != -INFINITY) S += bs[r * streams + s] * expf(m - M);
  }
  const float l = M + logf(S), base = run_score[(size_t)b * beams + j];
  if (lane == 0) lse[r] = l;
  for (int s = lane; s < n; s += 32) {
    const int c = tk_i[r * n + s];
    cv[s] = c >= 0 ? (tk_v[r * n + s] - l) + base : -INFINITY;  // log_softmax first, then + running score (HF :3252-3256,3283)
    ci[s] = c >= 0 ? j * V + c : 0x7fffffff;
  }
  __syncwarp();
  for (int out = 0; out < K; ++out) {
    float bv = -INFINITY; int bi = 0x7fffffff, bslot = -1;
    for (int s = lane; s < n; s += 32)
      if (bcand_better(cv[s], ci[s], bv, bi)) { bv = cv[s]; bi = ci[s]; bslot = s; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), os = __shfl_xor_sync(0xffffffffu, bslot, o);
      if (bcand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bslot = os; }
    }
    if (lane == 0) {
      row_val[r * K + out] = bv;
      row_idx[r * K + out] = bi;
      if (bslot >= 0) { cv[bslot] = -INFINITY; ci[bslot] = 0x7fffffff; }
    }
    __syncwarp();
  }
}

int launch_beam_topk_streams(const float* tk_v, const int* tk_i, const float* bm, const float* bs, int streams, int B, int rows_per_image, int n_live,
                             const float* run_score, int beams, int V, int K, float* lse, float* row_val, int* row_idx, float* cand_score, int* cand_idx,
                             cudaStream_t st) {
  GIC_REQUIRE(K == 2 * beams && beams >= 2 && beams <= 8 && n_live >= 1 && n_live <= beams && K <= GEMM_BEAM_KEEP && streams >= 1,
              "beam_topk_streams: beams %d / K %d / live %d / streams %d out of range", beams, K, n_live, streams);
  const size_t smem = (size_t)streams * GEMM_BEAM_KEEP * 8;
  GIC_REQUIRE(smem <= 48 * 1024, "beam_topk_streams: %d streams do not fit the merge kernel's shared memory", streams);
  beam_streams_kernel<<<B * n_live, 32, smem, st>>>(tk_v, tk_i, bm, bs, streams, rows_per_image, n_live, run_score, beams, V, K, lse, row_val, row_idx);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  beam_merge_kernel<<<B, 32, 0, st>>>(row_val, row_idx, rows_per_image, n_live, K, cand_score, cand_idx);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

int launch_beam_topk(const float* logits, int B, int rows_per_image, int n_live, const float* run_score, int beams, int V, int K,
                     float* lse, float* row_val, int* row_idx, float* cand_score, int* cand_idx, cudaStream_t st) {
  GIC_REQUIRE(K == 2 * beams && beams >= 2 && beams <= 8 && n_live >= 1 && n_live <= beams, "beam_topk: beams %d / K %d / live %d out of range", beams, K, n_live);
  beam_lse_kernel<<<B * n_live, 256, 0, st>>>(logits, rows_per_image, n_live, V, lse);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  switch (K) {
    case 4: GIC_TRY(launch_beam_rowtopk<4>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    case 6: GIC_TRY(launch_beam_rowtopk<6>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    case 8: GIC_TRY(launch_beam_rowtopk<8>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    case 10: GIC_TRY(launch_beam_rowtopk<10>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    case 12: GIC_TRY(launch_beam_rowtopk<12>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    case 14: GIC_TRY(launch_beam_rowtopk<14>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
    default: GIC_TRY(launch_beam_rowtopk<16>(logits, B, rows_per_image, n_live, run_score, lse, beams, V, row_val, row_idx, st)); break;
  }
  beam_merge_kernel<<<B, 32, 0, st>>>(row_val, row_idx, rows_per_image, n_live, K, cand_score, cand_idx);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// One thread per image: everything after the top-K of one HF beam-search iteration (steps c..g of :3285-3370).
__global__ void beam_update_kernel(BeamState s, int step, float len_denom /* (step+1)^length_penalty */) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  const int nb = s.beams, K = 2 * nb, N = s.max_new, V = s.V;
  constexpr int MAXB = 8, MAXK = 16;
  const int* old_run = s.run_seq[step & 1] + (size_t)b * nb * N;
  int* new_run = s.run_seq[(step + 1) & 1] + (size_t)b * nb * N;
  const int* old_fin = s.fin_seq[step & 1] + (size_t)b * nb * N;
  int* new_fin = s.fin_seq[(step + 1) & 1] + (size_t)b * nb * N;

  float cs[MAXK];
  int cbeam[MAXK], ctok[MAXK];
  bool cstop[MAXK];
  for (int j = 0; j < K; ++j) {
    cs[j] = s.cand_score[(size_t)b * K + j];
    const int flat = s.cand_idx[(size_t)b * K + j];
    cbeam[j] = flat / V;  // :2984-2987
    ctok[j] = flat % V;
    cstop[j] = (ctok[j] == s.eos) || (step + 1 >= N);  // EosTokenCriteria | MaxLengthCriteria
  }
  // ---- e. running beams for the next iteration: top `nb` of score + stop * -1e9 (candidates arrive sorted) ----
  int order[MAXK];
  int n = 0;
  for (int j = 0; j < K; ++j) if (!cstop[j]) order[n++] = j;
  for (int j = 0; j < K; ++j) if (cstop[j]) order[n++] = j;
  for (int r = 0; r < nb; ++r) {
    const int j = order[r];
    for (int t = 0; t < N; ++t) new_run[r * N + t] = (t == step) ? ctok[j] : old_run[cbeam[j] * N + t];
    s.run_score[(size_t)b * nb + r] = cstop[j] ? cs[j] + (-1.0e9f) : cs[j];
    s.beam_idx[(size_t)b * nb + r] = b * nb + cbeam[j];
    s.next_tok[(size_t)b * nb + r] = ctok[j];
  }
  // ---- f. finished pool ----
  const bool unsat = s.unsat[b] != 0;
  float ms[MAXB + MAXK];
  for (int i = 0; i < nb; ++i) ms[i] = s.fin_score[(size_t)b * nb + i];
  for (int j = 0; j < K; ++j) {
    const bool did = cstop[j] && j < nb;  // only the top `nb` candidates may be finalised (:3045)
    float v = cs[j] / len_denom;
    v += unsat ? 0.f : -1.0e9f;
    v += did ? 0.f : -1.0e9f;
    ms[nb + j] = v;
  }
  bool used[MAXB + MAXK];
  for (int i = 0; i < nb + K; ++i) used[i] = false;
  float nscore[MAXB];
  unsigned char nflag[MAXB];
  int nlen[MAXB];
  for (int r = 0; r < nb; ++r) {
    int best = -1;
    for (int i = 0; i < nb + K; ++i)
      if (!used[i] && (best < 0 || ms[i] > ms[best])) best = i;
    used[best] = true;
    nscore[r] = ms[best];
    if (best < nb) {
      for (int t = 0; t < N; ++t) new_fin[r * N + t] = old_fin[best * N + t];
      nflag[r] = s.fin_flag[(size_t)b * nb + best];
      nlen[r] = s.fin_len[(size_t)b * nb + best];
    } else {
      const int j = best - nb;
      for (int t = 0; t < N; ++t) new_fin[r * N + t] = (t == step) ? ctok[j] : old_run[cbeam[j] * N + t];
      nflag[r] = (cstop[j] && j < nb) ? 1 : 0;
      nlen[r] = step + 1;
    }
  }
  float min_fin = INFINITY;
  for (int r = 0; r < nb; ++r) {
    s.fin_score[(size_t)b * nb + r] = nscore[r];
    s.fin_flag[(size_t)b * nb + r] = nflag[r];
    s.fin_len[(size_t)b * nb + r] = nlen[r];
    min_fin = fminf(min_fin, nscore[r]);
  }
  // ---- early-stop heuristic (:2876-2921, early_stopping=False): can the best running beam still beat the worst finished? ----
  const float best_running = s.run_score[(size_t)b * nb] / len_denom;  // cur_len after the increment = step + 1
  bool improve = false;
  for (int r = 0; r < nb; ++r) improve |= best_running > (nflag[r] ? min_fin : -1.0e9f);
  s.unsat[b] = (unsat && improve) ? 1 : 0;
}

int launch_beam_update(const BeamState& s, int step, float len_denom, cudaStream_t st) {
  GIC_REQUIRE(s.beams >= 1 && s.beams <= 8, "beam_update: beams %d out of range (<= 8)", s.beams);
  beam_update_kernel<<<ceil_div(s.B, 64), 64, 0, st>>>(s, step, len_denom);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

__global__ void beam_init_kernel(BeamState s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = s.beams, N = s.max_new;
  if (i < s.B * nb * N) {
    s.run_seq[0][i] = s.eos; s.run_seq[1][i] = s.eos;  // output_fill_value = pad = eos (:3188-3196)
    s.fin_seq[0][i] = s.eos; s.fin_seq[1][i] = s.eos;
  }
  if (i < s.B * nb) {
    s.run_score[i] = (i % nb == 0) ? 0.f : -1.0e9f;  // only beam 0 is live at the first step (:3200-3201)
    s.fin_score[i] = -1.0e9f;
    s.fin_flag[i] = 0;
    s.fin_len[i] = 0;
    s.beam_idx[i] = (i / nb) * nb;
    s.next_tok[i] = s.eos;
  }
  if (i < s.B) s.unsat[i] = 1;
}

int launch_beam_init(const BeamState& s, cudaStream_t st) {
  const int n = s.B * s.beams * s.max_new;
  beam_init_kernel<<<ceil_div(n, 256), 256, 0, st>>>(s);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// next decode input of every beam row: h[row] = wte[next_tok[row]] + wpe[pos]
__global__ void __launch_bounds__(128) beam_embed_kernel(const int* __restrict__ next_tok, const float* __restrict__ wte_f32,
                                                         const bf16* __restrict__ wte_bf16, const float* __restrict__ wpe, int pos, int d,
                                                         float* __restrict__ h) {
  const int row = blockIdx.x;
  const int tok = next_tok[row];
  const float* pe = wpe + (size_t)pos * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float e = wte_f32 ? wte_f32[(size_t)tok * d + c] : __bfloat162float(wte_bf16[(size_t)tok * d + c]);
    h[(size_t)row * d + c] = e + pe[c];
  }
}

__global__ void beam_ancestry_kernel(const int* __restrict__ anc_old, int* __restrict__ anc_new, const int* __restrict__ beam_idx, int rows, int ld, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * t) return;
  const int r = i / t, g = i - r * t;
  const int p = beam_idx[r];
  anc_new[(size_t)r * ld + g] = (g == t - 1) ? p : anc_old[(size_t)p * ld + g];
}

int launch_beam_ancestry(const int* anc_old, int* anc_new, const int* beam_idx, int rows, int ld, int t, cudaStream_t st) {
  if (t <= 0) return GIC_OK;
  beam_ancestry_kernel<<<ceil_div(rows * t, 256), 256, 0, st>>>(anc_old, anc_new, beam_idx, rows, ld, t);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

int launch_beam_embed(const int* next_tok, const float* wte_f32, const bf16* wte_bf16, const float* wpe, int pos, int d, float* h, int rows,
                      cudaStream_t st) {
  beam_embed_kernel<<<rows, 128, 0, st>>>(next_tok, wte_f32, wte_bf16, wpe, pos, d, h);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }
int launch_set_int(int* p, int v, cudaStream_t st) {
  set_int_kernel<<<1, 1, 0, st>>>(p, v);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// best finished hypothesis of every image -> ids_out [B, N] (eos padded), score, and the longest selected length
__global__ void beam_finalize_kernel(BeamState s, int final_buf, int64_t* __restrict__ ids_out, float* __restrict__ scores_out,
                                     int* __restrict__ gen_len_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  const int nb = s.beams, N = s.max_new;
  const int* seq = s.fin_seq[final_buf] + (size_t)b * nb * N;  // slot 0 = best (pool kept sorted by score)
  for (int t = 0; t < N; ++t) ids_out[(size_t)b * N + t] = seq[t];
  if (scores_out) scores_out[b] = s.fin_score[(size_t)b * nb];
  if (gen_len_out) atomicMax(gen_len_out, s.fin_len[(size_t)b * nb]);
}

int launch_beam_finalize(const BeamState& s, int final_buf, int64_t* ids_out, float* scores_out, int* gen_len_out, cudaStream_t st) {
  if (gen_len_out) GIC_TRY(launch_set_int(gen_len_out, 0, st));
  beam_finalize_kernel<<<ceil_div(s.B, 128), 128, 0, st>>>(s, final_buf, ids_out, scores_out, gen_len_out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

}  // namespace gic
