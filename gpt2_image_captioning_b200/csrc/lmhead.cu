// Token selection: argmax over the LM-head logits and the per-step bookkeeping of the greedy loop.
// Replaces src/models.py:398-469 per step: logits[:, -1, :] -> argmax (ties -> LOWEST index, as torch.argmax) ->
// finished |= (tok == eos); tok[finished] = eos -> append -> wte(tok) as the next input (+ wpe of its position,
// HF:models/gpt2/modeling_gpt2.py:579-585).  Everything stays on the device: no per-step host sync
// (the reference does `is_finished.all()` on the host every step, src/models.py:390).
#include <atomic>

#include "kernels.cuh"

namespace gic {


__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void block_argmax(float& v, int& idx, float* sv, int* si) {
  // warp reduce, then across warps; keeps the lowest index among equal maxima
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (better(ov, oi, v, idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { sv[warp] = v; si[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sv[lane] : -INFINITY;
    idx = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (better(ov, oi, v, idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { sv[0] = v; si[0] = idx; }
  }
  __syncthreads();
  v = sv[0];
  idx = si[0];
}

// fp32 mode: logits [B,V] were materialised by the CUDA-core GEMM; scan them in LMHEAD_F32_PARTS column slabs per row.
__global__ void __launch_bounds__(256) argmax_partials_kernel(const float* logits, int B, int V, float* part_val, int* part_idx, int part_ld) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int b = blockIdx.x, part = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const int chunk = (V + LMHEAD_F32_PARTS - 1) / LMHEAD_F32_PARTS;
  const int c0 = part * chunk, c1 = min(V, c0 + chunk);
  float v = -INFINITY;
  int idx = 0x7fffffff;
  const float* row = logits + (size_t)b * V;
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const float x = __ldcg(row + c);  // produced by the previous kernel (not an invariant load)
    if (better(x, c, v, idx)) { v = x; idx = c; }
  }
  block_argmax(v, idx, sv, si);
  if (threadIdx.x == 0) {
    part_val[(size_t)b * part_ld + part] = v;
    part_idx[(size_t)b * part_ld + part] = idx;
  }
}

int launch_argmax_partials(const float* logits, int B, int V, float* part_val, int* part_idx, int part_ld, cudaStream_t st) {
  dim3 grid(B, LMHEAD_F32_PARTS);
  GIC_CHECK_CUDA(launch_kernel(argmax_partials_kernel, grid, dim3(256), 0, st, logits, B, V, part_val, part_idx, part_ld));
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Temperature / nucleus (top-p) sampling: the `temperature > 0` branch of ImageCaptioningModel.generate
// (src/models.py:400-449).  One block per row; the row's V logits arrive in shared memory (50 257 x 4 B = 196 KB) as eight
// cp.async.bulk chunks (the running maximum is taken chunk by chunk while the later ones are in flight) and are replaced in place by
// the unnormalised probabilities p_i = exp((z_i - max) / T).  The reference sorts the row, takes the cumulative softmax and keeps
// every token up to and including the first whose cumulative probability exceeds top_p (:413-432), then draws from the renormalised
// kept set.  That kept set is {i : sum of p_j over p_j > p_i  <=  top_p S}, so drawing from it is REJECTION SAMPLING on the full
// distribution: draw a token by inverse CDF in index order (one Philox4x32-10 uniform per (seed, row, step, attempt)), accept it if
// the mass strictly above its probability is <= top_p S (one block-wide sum), else draw again.  The acceptance probability is the
// nucleus mass >= top_p, so 1 / top_p attempts on average (two sweeps of shared memory each) instead of a sort or a 31-pass
// bisection.  After SAMPLE_ATTEMPTS rejections (tiny top_p on a flat row) the exact threshold is found by bisection over the float
// bit patterns and the token drawn from the kept mass directly -- the same distribution, so the mixture is exact.
// Parity with torch.multinomial is distributional only.  The token is handed to finalize_token_kernel as a single (value, index)
// "partial", so the EOS rules are shared with greedy.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t c0, uint32_t c1, uint32_t c2 = 0u) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, h1;
    const uint32_t l0 = mulhilo32(0xD2511F53u, x0, &h0), l1 = mulhilo32(0xCD9E8D57u, x2, &h1);
    const uint32_t y0 = h1 ^ x1 ^ k0, y1 = l1, y2 = h0 ^ x3 ^ k1, y3 = l0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(x0 >> 8) * (1.0f / 16777216.0f);  // [0, 1)
}

constexpr int SAMPLE_THREADS = 512;
constexpr int SAMPLE_CHUNKS = 8;    // bulk copies per row
constexpr int SAMPLE_ATTEMPTS = 4;  // rejection rounds before the exact bisection
constexpr int SAMPLE_PAD = 8;       // floats of slack in front of / behind the row in shared memory (16-byte alignment of the bulk part)

__device__ __forceinline__ float sample_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // `red` may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_top_p_kernel(const float* __restrict__ logits, int V, long ld, float inv_temperature, float top_p,
                                                                      unsigned long long seed, const int* __restrict__ d_step, int step_override,
                                                                      float* __restrict__ part_val, int* __restrict__ part_idx, int part_ld,
                                                                      const SampleParams* __restrict__ dev_params) {
  extern __shared__ __align__(16) float sample_smem[];  // [SAMPLE_PAD + V + SAMPLE_PAD]
  __shared__ float red[SAMPLE_THREADS / 32];
  __shared__ float s_scan[SAMPLE_THREADS];
  __shared__ float s_wtot[SAMPLE_THREADS / 32];
  __shared__ int s_tok;
  __shared__ float s_ptok;
  __shared__ __align__(8) unsigned long long s_bar[SAMPLE_CHUNKS];
  if (dev_params) {  // graph replays: the call's temperature / top_p / seed live in device memory, not in the captured arguments
    inv_temperature = __ldcg(&dev_params->inv_temperature);
    top_p = __ldcg(&dev_params->top_p);
    seed = __ldcg(&dev_params->seed);
  }
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* z = logits + (size_t)b * ld;
  // the row in three parts: `head` floats up to the first 16-byte boundary of the global address, `body` floats in whole 16-byte
  // units (bulk copies), `tail` floats after them; prob[i] sits where the body lands 16-byte aligned in shared memory too
  const int head = min(V, (int)(((16u - (unsigned)((uintptr_t)z & 15u)) & 15u) >> 2));
  const int body = ((V - head) >> 2) << 2;
  float* prob = sample_smem + SAMPLE_PAD - head;
  const int per = ((((body + SAMPLE_CHUNKS - 1) / SAMPLE_CHUNKS) + 3) >> 2) << 2;  // floats per chunk (multiple of 4)
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
  if (t == 0) {
    for (int c = 0; c < SAMPLE_CHUNKS; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * c) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int c = 0; c < SAMPLE_CHUNKS; ++c) {
      const int c0 = min(body, c * per), c1 = min(body, c0 + per);
      const uint32_t bytes = (uint32_t)(c1 - c0) * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * c), "r"(bytes) : "memory");
      if (bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(prob + head + c0)),
                     "l"(z + head + c0), "r"(bytes), "r"(bar0 + 8 * c)
                     : "memory");
    }
  }
  // the unaligned ends (at most 3 + 3 floats) through ordinary loads
  float m = -INFINITY;
  if (t < head) { const float v = z[t]; prob[t] = v; m = v; }
  if (head + body + t < V) { const float v = z[head + body + t]; prob[head + body + t] = v; m = fmaxf(m, v); }
  __syncthreads();  // barriers initialised before anyone polls them
  for (int c = 0; c < SAMPLE_CHUNKS; ++c) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar0 + 8 * c), "r"(0u) : "memory");
    const int c0 = head + min(body, c * per), c1 = head + min(body, c * per + per);
    for (int i = c0 + t; i < c1; i += SAMPLE_THREADS) m = fmaxf(m, prob[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) m = fmaxf(m, red[w]);
  // p_i in place; thread t owns the contiguous slice [t * seg, (t + 1) * seg) (bank = (3 t + j) mod 32: conflict-free) -- the
  // inverse CDF below walks the row in index order
  const int seg = (V + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
  const int i0 = min(V, t * seg), i1 = min(V, i0 + seg);
  float part = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float p = expf((prob[i] - m) * inv_temperature);  // softmax(z / T) up to the common factor
    prob[i] = p;
    part += p;
  }
  // inclusive scan of the 512 slice sums (fixed order: deterministic)
  float incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();  // (also: every p_i is in place)
  float woff = 0.f, total = 0.f;
#pragma unroll
  for (int w = 0; w < SAMPLE_THREADS / 32; ++w) {
    if (w < warp) woff += s_wtot[w];
    total += s_wtot[w];
  }
  s_scan[t] = woff + incl;
  const float before = woff + incl - part;
  const int step = step_override >= 0 ? step_override : __ldcg(d_step);
  const float want = top_p * total;
  // token at inverse-CDF position `target` of the mass selected by `keep` (bit pattern threshold; 0 = the whole row)
  auto draw = [&](float target, float my_before, float my_after, float my_part, uint32_t keep) {
    if (t == 0) s_tok = -1;
    __syncthreads();
    if (my_part > 0.f && target >= my_before && target < my_after) {
      float run = my_before;
      int tok = -1;
      float ptok = 0.f;
      for (int i = i0; i < i1; ++i) {
        const float p = prob[i];
        if (p > 0.f && __float_as_uint(p) >= keep) {
          tok = i;  // the last kept token of the slice catches rounding at its upper edge
          ptok = p;
          run += p;
          if (target < run) break;
        }
      }
      s_tok = tok;
      s_ptok = ptok;
    }
    __syncthreads();
    if (s_tok < 0) {  // target landed on / beyond the total through rounding: the last kept token of the row
      if (t == 0) {
        for (int i = V - 1; i >= 0; --i)
          if (prob[i] > 0.f && __float_as_uint(prob[i]) >= keep) { s_tok = i; s_ptok = prob[i]; break; }
      }
      __syncthreads();
    }
  };
  bool accepted = false;
  for (int attempt = 0; attempt < SAMPLE_ATTEMPTS && !accepted; ++attempt) {
    const float target = philox_uniform(seed, (uint32_t)b, (uint32_t)step, (uint32_t)attempt) * total;
    draw(target, before, s_scan[t], part, 0u);
    if (top_p >= 1.0f) { accepted = true; break; }
    const float ptok = s_ptok;
    float above = 0.f;
    for (int i = t; i < V; i += SAMPLE_THREADS) {
      const float p = prob[i];
      above += p > ptok ? p : 0.f;
    }
    accepted = sample_block_sum(above, red) <= want;  // (block-uniform: every thread holds the same sum)
  }
  if (!accepted) {
    // exact path.  tau: bit pattern of the smallest kept probability.  Invariant: tail(lo) > top_p * S >= tail(hi)   (tail(x) = sum of p >= x)
    uint32_t lo = 0u, hi = __float_as_uint(1.0f) + 1u;
    while (hi - lo > 1u) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      float a = 0.f;
      for (int i = t; i < V; i += SAMPLE_THREADS) {
        const float p = prob[i];
        a += (__float_as_uint(p) >= mid) ? p : 0.f;
      }
      if (sample_block_sum(a, red) > want) lo = mid; else hi = mid;
    }
    float kpart = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float p = prob[i];
      kpart += (__float_as_uint(p) >= lo) ? p : 0.f;
    }
    __syncthreads();
    s_scan[t] = kpart;
    __syncthreads();
    if (t == 0) {  // serial inclusive scan of 512 partials (rare path)
      float run = 0.f;
      for (int j = 0; j < SAMPLE_THREADS; ++j) { run += s_scan[j]; s_scan[j] = run; }
    }
    __syncthreads();
    const float ktotal = s_scan[SAMPLE_THREADS - 1];
    const float target = philox_uniform(seed, (uint32_t)b, (uint32_t)step, (uint32_t)SAMPLE_ATTEMPTS) * ktotal;
    draw(target, t == 0 ? 0.f : s_scan[t - 1], s_scan[t], kpart, lo);
  }
  if (t == 0) {
    part_val[(size_t)b * part_ld] = 1.0f;
    part_idx[(size_t)b * part_ld] = s_tok;
  }
}

__global__ void set_sample_params_kernel(SampleParams* dst, float inv_temperature, float top_p, unsigned long long seed) {
  dst->inv_temperature = inv_temperature;
  dst->top_p = top_p;
  dst->seed = seed;
}

int launch_set_sample_params(SampleParams* dst, float temperature, float top_p, unsigned long long seed, cudaStream_t st) {
  GIC_REQUIRE(dst != nullptr && temperature > 0.f && top_p > 0.f, "set_sample_params: bad argument");
  set_sample_params_kernel<<<1, 1, 0, st>>>(dst, 1.0f / temperature, top_p, seed);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

int launch_sample_top_p(const float* logits, int B, int V, float temperature, float top_p, unsigned long long seed, const int* d_step,
                        int step_override, float* part_val, int* part_idx, int part_ld, cudaStream_t st, const SampleParams* dev_params, long ld) {
  GIC_REQUIRE(temperature > 0.f, "sample_top_p: temperature must be > 0 (0 is the greedy path)");
  GIC_REQUIRE(top_p > 0.f, "sample_top_p: top_p must be > 0");
  const size_t smem = ((size_t)V + 2 * SAMPLE_PAD) * sizeof(float);
  GIC_REQUIRE(smem <= 200 * 1024, "sample_top_p: a row of %d probabilities does not fit in shared memory", V);
  if (ld <= 0) ld = V;
  GIC_REQUIRE(ld >= V, "sample_top_p: leading dimension %ld < V = %d", ld, V);
  static std::atomic<bool> configured{false};  // (engine contexts may be driven from several host threads)
  if (!configured.load(std::memory_order_acquire)) {
    GIC_CHECK_CUDA(cudaFuncSetAttribute(sample_top_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(true, std::memory_order_release);
  }
  sample_top_p_kernel<<<B, SAMPLE_THREADS, smem, st>>>(logits, V, ld, 1.0f / temperature, top_p, seed, d_step, step_override, part_val, part_idx, part_ld,
                                                       dev_params);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact greedy token from an approximate LM head (bf16x2 engine).  The head runs ONCE on the tensor cores with single bf16
// operands (a_hi . wte_hi, 1 MMA per product instead of 3) and its epilogue keeps, per (row, slot of 32..128 columns), the best
// value, its column and the runner-up value.  Rounding both operands to bf16 moves logit n by at most
//   err_n <= 2^-8 (1 + 2^-10) sum_k |a_k w_nk| (+ fp32 accumulation)  <=  DELTA_REL |a|_2 |wte_n|_2        (Cauchy-Schwarz),
// so with i* = the approximate argmax, the true argmax n satisfies  approx_n + err_n >= approx_i* - err_i*.  Per row:
//   lm_head_candidates_kernel  recomputes a = ln_f(h) in fp32 (HF:models/gpt2/modeling_gpt2.py:628), finds i*, and walks the
//       slots: a slot whose best value plus the slot's error bound (largest |wte_n| of its columns, precomputed per tiling)
//       reaches the lower bound contributes its best column -- or all of its columns when its RUNNER-UP reaches it too -- to a
//       global (row, column) list;
//   lm_head_rescore_pairs_kernel  re-scores the listed pairs with fp32 dot products a . wte_f32[n] (:705-706, tied head), one warp
//       per pair over the whole GPU (load-balanced: a row with a flat distribution does not become a straggler), and keeps the
//       row's best in a packed (value, ~column) key: largest value, then LOWEST index (torch.argmax, src/models.py:441);
//       rows whose candidates did not fit the list are scanned over all V columns by the whole grid.
// finalize_token_kernel unpacks the key.  The bound is rigorous, so the token is the argmax of the fp32 logits of ln_f(h).
// ---------------------------------------------------------------------------------------------------------------
constexpr float RESCORE_DELTA_REL = 0.00390625f * 1.02f;  // 2^-8 with 2 % slack for the 2^-18 cross term and the fp32 accumulation
constexpr int RESCORE_THREADS = 128;

__device__ __forceinline__ unsigned long long pack_best(float v, int idx) {
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of fp32 onto unsigned
  return ((unsigned long long)u << 32) | (unsigned int)(0x7fffffff - idx);
}

__global__ void __launch_bounds__(RESCORE_THREADS) lm_head_candidates_kernel(RescoreArgs r) {
  extern __shared__ float rs_a[];  // [d] ln_f(h) in fp32
  __shared__ float red[RESCORE_THREADS / 32];
  __shared__ float s_bv[RESCORE_THREADS / 32];
  __shared__ int s_bi[RESCORE_THREADS / 32];
  __shared__ int s_row_n;
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (r.row_map && __ldcg(r.row_map + b) < 0) return;  // padding slot of a compacted batch (its activations are garbage)
  const int tslot = trace_begin(r.step_trace, TRACE_RESCORE, 0);
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
  };
  const int d = r.d;
  // ---- ln_f(h) in fp32, two-pass like layernorm_kernel; kept in global memory for the pair kernel ----
  const float* hr = r.h + (size_t)b * r.h_row_stride;
  float s = 0.f;
  for (int c = t; c < d; c += RESCORE_THREADS) { const float v = __ldcg(hr + c); rs_a[c] = v; s += v; }
  const float mean = block_sum(s) / (float)d;
  float q = 0.f;
  for (int c = t; c < d; c += RESCORE_THREADS) { const float v = rs_a[c] - mean; q += v * v; }
  const float rstd = 1.0f / sqrtf(block_sum(q) / (float)d + 1e-5f);
  float n2 = 0.f;
  float* ag = r.a_f32 + (size_t)b * d;
  for (int c = t; c < d; c += RESCORE_THREADS) {
    const float v = (rs_a[c] - mean) * rstd * __ldg(r.lnw + c) + __ldg(r.lnb + c);
    ag[c] = v;
    n2 += v * v;
  }
  const float dn = RESCORE_DELTA_REL * sqrtf(block_sum(n2));  // error bound per unit of |wte_n|
  // ---- approximate argmax over the row's slots ----
  const float* pv = r.part_val + (size_t)b * r.part_ld;
  const int* pi = r.part_idx + (size_t)b * r.part_ld;
  const float* pv2 = r.part_val2 + (size_t)b * r.part_ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int p = t; p < r.n_parts; p += RESCORE_THREADS) {
    const float v = __ldcg(pv + p);
    const int i = __ldcg(pi + p);
    if (better(v, i, m, mi)) { m = v; mi = i; }
  }
  if (t == 0) { s_row_n = 0; r.row_flag[b] = 0; }
  block_argmax(m, mi, s_bv, s_bi);
  if (t == 0) r.best[b] = pack_best(-INFINITY, 0x7fffffff);
  if (mi < 0 || mi >= r.V) return;  // (a row of NaNs: nothing to re-score; finalize sees the sentinel)
  const float lb = m - dn * __ldg(r.wte_norm + mi);
  // ---- candidates ----
  const int block_n = r.block_n;
  for (int p = t; p < r.n_parts; p += RESCORE_THREADS) {
    const float e = dn * __ldg(r.slot_norm_max + p);
    if (__ldcg(pv + p) + e < lb) continue;
    const bool rescan = __ldcg(pv2 + p) + e >= lb;
    const int n0 = (p >> 1) * block_n + (p & 1) * 32, n_end = min((p >> 1) * block_n + block_n, r.V);
    int k = 1;
    if (rescan) { k = 0; for (int c0 = n0; c0 < n_end; c0 += 64) k += max(0, min(32, n_end - c0)); }
    const int mine = atomicAdd(&s_row_n, k);
    const int base = (mine + k <= r.row_budget) ? atomicAdd(r.pair_count, k) : r.pair_cap;
    if (base + k > r.pair_cap) {  // no room (pathologically flat row / full list): the whole row is scanned by the pair kernel
      if (atomicExch(r.row_flag + b, 1) == 0) r.flag_rows[atomicAdd(r.flag_count, 1)] = b;
      continue;
    }
    if (!rescan) r.pairs[base] = make_int2(b, __ldcg(pi + p));
    else {
      int o = base;
      for (int c0 = n0; c0 < n_end; c0 += 64)
        for (int c = c0; c < c0 + 32 && c < n_end; ++c) r.pairs[o++] = make_int2(b, c);
    }
  }
  trace_end(r.step_trace, tslot);
}

// exact fp32 logit of (row, col): a[row] . wte[col], lanes over k with 128-bit loads, fixed summation order
__device__ __forceinline__ float rescore_dot(const float* __restrict__ a, const float* __restrict__ w, int d, int lane) {
  float acc = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 av = __ldcg(reinterpret_cast<const float4*>(a + c));
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
    acc = fmaf(av.x, wv.x, acc); acc = fmaf(av.y, wv.y, acc); acc = fmaf(av.z, wv.z, acc); acc = fmaf(av.w, wv.w, acc);
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(RESCORE_THREADS) lm_head_rescore_pairs_kernel(RescoreArgs r) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * RESCORE_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * RESCORE_THREADS) >> 5;
  pdl_launch_dependents();
  pdl_wait();
  const int tslot = trace_begin(r.step_trace, TRACE_RESCORE, 1);
  const int n = min(__ldcg(r.pair_count), r.pair_cap);
  // two pairs per warp and iteration: their loads overlap
  for (int i = gw; i < n; i += 2 * nw) {
    const int2 p0 = __ldcg(r.pairs + i);
    const bool two = i + nw < n;
    const int2 p1 = two ? __ldcg(r.pairs + i + nw) : p0;
    const float v0 = rescore_dot(r.a_f32 + (size_t)p0.x * r.d, r.wte + (size_t)p0.y * r.d, r.d, lane);
    const float v1 = rescore_dot(r.a_f32 + (size_t)p1.x * r.d, r.wte + (size_t)p1.y * r.d, r.d, lane);
    if (lane == 0) {
      atomicMax(r.best + p0.x, pack_best(v0, p0.y));
      if (two) atomicMax(r.best + p1.x, pack_best(v1, p1.y));
    }
  }
  // rows that did not fit the list: every column, the whole grid
  const int nf = __ldcg(r.flag_count);
  for (int f = 0; f < nf; ++f) {
    const int row = __ldcg(r.flag_rows + f);
    unsigned long long bk = 0ull;
    for (int col = gw; col < r.V; col += nw) {
      const float v = rescore_dot(r.a_f32 + (size_t)row * r.d, r.wte + (size_t)col * r.d, r.d, lane);
      const unsigned long long k = pack_best(v, col);
      bk = k > bk ? k : bk;
    }
    if (lane == 0 && bk) atomicMax(r.best + row, bk);
  }
  trace_end(r.step_trace, tslot);
}

int launch_lm_head_rescore(const RescoreArgs& r0, cudaStream_t st) {
  RescoreArgs r = r0;
  GIC_REQUIRE(r.d % 4 == 0 && r.d <= 4096 && r.n_parts >= 1 && r.part_ld >= r.n_parts && r.rows > 0, "lm_head_rescore: bad sizes d=%d n_parts=%d part_ld=%d", r.d,
              r.n_parts, r.part_ld);
  r.step_trace = trace_desc();
  GIC_CHECK_CUDA(launch_kernel(lm_head_candidates_kernel, dim3(r.rows), dim3(RESCORE_THREADS), (size_t)r.d * sizeof(float), st, r));
  note_launch();
  r.step_trace = trace_desc();
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  GIC_CHECK_CUDA(launch_kernel(lm_head_rescore_pairs_kernel, dim3(sms * 4), dim3(RESCORE_THREADS), 0, st, r));
  note_launch();
  return GIC_OK;
}

// per-row norms of the fp32 embedding table and, for one LM-head tiling, the largest norm inside each (tile, column-parity) slot
__global__ void __launch_bounds__(128) row_norm_kernel(const float* __restrict__ w, int N, int K, float* __restrict__ out) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) { const float v = w[(size_t)n * K + k]; s += v * v; }
  s = warp_sum(s);
  if (lane == 0) out[n] = sqrtf(s);
}
__global__ void slot_norm_max_kernel(const float* __restrict__ norms, int V, int block_n, float* __restrict__ out, int n_slots) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_slots) return;
  const int n0 = (p >> 1) * block_n + (p & 1) * 32, n_end = min((p >> 1) * block_n + block_n, V);
  float m = 0.f;
  for (int c0 = n0; c0 < n_end; c0 += 64)
    for (int c = c0; c < c0 + 32 && c < n_end; ++c) m = fmaxf(m, norms[c]);
  out[p] = m;
}
int launch_row_norms(const float* w, int N, int K, float* out, cudaStream_t st) {
  row_norm_kernel<<<ceil_div(N, 4), 128, 0, st>>>(w, N, K, out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}
int launch_slot_norm_max(const float* norms, int V, int block_n, float* out, cudaStream_t st) {
  const int n_slots = 2 * ceil_div(V, block_n);
  slot_norm_max_kernel<<<ceil_div(n_slots, 128), 128, 0, st>>>(norms, V, block_n, out, n_slots);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// One block per row: reduce the partial maxima, apply the EOS rules, record the token and build the next input.
__global__ void __launch_bounds__(128) finalize_token_kernel(FinalizeArgs a) {
  __shared__ float sv[4];
  __shared__ int si[4];
  __shared__ int s_tok;
  const int b = blockIdx.x;  // activation slot; its caption is row `orig` of the call (compacted batches: row_map, -1 = padding slot)
  pdl_launch_dependents();
  pdl_wait();
  const int tslot = trace_begin(a.step_trace, TRACE_FINALIZE, 0);
  const int step = __ldcg(a.d_step);
  const int orig = a.row_map ? __ldcg(a.row_map + b) : b;
  const bool dead = orig < 0;  // block-uniform
  float v = -INFINITY;
  int idx = 0x7fffffff;
  if (a.packed_best && !dead && threadIdx.x == 0) {  // exactly re-scored head: the row's (value, ~column) key
    const unsigned long long k = __ldcg(a.packed_best + b);
    idx = 0x7fffffff - (int)(unsigned int)(k & 0xffffffffull);
    v = 0.f;
  }
  for (int p = threadIdx.x; p < ((dead || a.packed_best) ? 0 : a.n_parts); p += blockDim.x) {
    const float pv = __ldcg(a.part_val + (size_t)b * a.part_ld + p);
    const int pi = __ldcg(a.part_idx + (size_t)b * a.part_ld + p);
    if (better(pv, pi, v, idx)) { v = pv; idx = pi; }
  }
  block_argmax(v, idx, sv, si);
  if (threadIdx.x == 0) {
    int tok = a.eos;
    unsigned char fin = 1;
    if (!dead) {
      tok = idx;
      fin = a.finished[orig];
      if (tok == a.eos && !fin) {  // src/models.py:453-455
        fin = 1;
        a.finished[orig] = 1;
        a.first_eos[orig] = step;
      }
      if (fin) tok = a.eos;  // :458-460
      a.ids_out[(size_t)orig * a.max_new + step] = (int64_t)tok;
    }
    s_tok = tok;
    if (fin && a.fin_counter) atomicAdd(a.fin_counter, 1);  // slots finished after this step (for the host's early exit / compaction)
  }
  __syncthreads();
  const int tok = s_tok;
  const int pos = a.P + step;  // position of this token when it is fed back
  float ssum = 0.f, ssq = 0.f;
  if (pos < a.n_pos && !dead) {
    const float* pe = a.wpe + (size_t)pos * a.d;
    float* hn = a.h_next + (size_t)b * a.d;
    if (a.wte_bf16 && a.d % 8 == 0) {
      // bf16 table: one 16-byte load of the embedding + two of the position row per thread, one round trip for the whole row
      for (int c = threadIdx.x * 8; c < a.d; c += blockDim.x * 8) {
        const uint4 raw = *reinterpret_cast<const uint4*>(a.wte_bf16 + (size_t)tok * a.d + c);
        const float4 p0 = *reinterpret_cast<const float4*>(pe + c), p1 = *reinterpret_cast<const float4*>(pe + c + 4);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float2 e0 = __bfloat1622float2(h2[0]), e1 = __bfloat1622float2(h2[1]), e2 = __bfloat1622float2(h2[2]), e3 = __bfloat1622float2(h2[3]);
        const float x[8] = {e0.x + p0.x, e0.y + p0.y, e1.x + p0.z, e1.y + p0.w, e2.x + p1.x, e2.y + p1.y, e3.x + p1.z, e3.y + p1.w};
        *reinterpret_cast<float4*>(hn + c) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(hn + c + 4) = make_float4(x[4], x[5], x[6], x[7]);
        if (a.hb_next) {
          uint4 ob;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ob);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            o2[u] = __floats2bfloat162_rn(x[2 * u], x[2 * u + 1]);
            const float2 xf = __bfloat1622float2(o2[u]);
            ssum += xf.x + xf.y;
            ssq += xf.x * xf.x + xf.y * xf.y;
          }
          *reinterpret_cast<uint4*>(a.hb_next + (size_t)b * a.d + c) = ob;
        }
      }
    } else {
      for (int c = threadIdx.x; c < a.d; c += blockDim.x) {
        const float x = (a.wte_f32 ? a.wte_f32[(size_t)tok * a.d + c] : __bfloat162float(a.wte_bf16[(size_t)tok * a.d + c])) + pe[c];
        hn[c] = x;
        if (a.hb_next) {
          const bf16 xb = __float2bfloat16_rn(x);
          a.hb_next[(size_t)b * a.d + c] = xb;
          float xf = __bfloat162float(xb);
          if (a.hb_next_lo) {  // bf16x2: hi + lo operand, statistics of the sum
            const bf16 xl = __float2bfloat16_rn(x - xf);
            a.hb_next_lo[(size_t)b * a.d + c] = xl;
            xf += __bfloat162float(xl);
          }
          ssum += xf;
          ssq += xf * xf;
        }
      }
    }
  }
  if (a.stats_next) {  // block-uniform
    __shared__ float s_sum[4], s_sq[4];
    ssum = warp_sum(ssum);
    ssq = warp_sum(ssq);
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = ssum; s_sq[threadIdx.x >> 5] = ssq; }
    __syncthreads();
    if (threadIdx.x == 0) a.stats_next[b] = make_float2((s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]), (s_sq[0] + s_sq[1]) + (s_sq[2] + s_sq[3]));
  }
  // the last block to finish advances the device-side step / position counters (all blocks have read them by then)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(a.done_counter, 1);
    if (t == (int)gridDim.x - 1) {
      *a.done_counter = 0;
      *a.d_pos = pos;
      *a.d_step = step + 1;
      if (a.rescore_counters) { a.rescore_counters[0] = 0; a.rescore_counters[1] = 0; }  // pair / flagged-row lists of the next step start empty
      if (a.fin_counter) {  // every row has emitted EOS: the reference loop would stop before the next step (src/models.py:390-391)
        const int nf = atomicExch(a.fin_counter, 0);
        if (nf == (int)gridDim.x) *a.all_done = 1;
        if (a.live_rows) *a.live_rows = (int)gridDim.x - nf;  // unfinished rows: the host shrinks the batch to them (compact_rows_kernel)
      }
    }
  }
  trace_end(a.step_trace, tslot);
}

int launch_finalize_token(const FinalizeArgs& a0, cudaStream_t st) {
  FinalizeArgs a = a0;
  a.step_trace = trace_desc();
  GIC_CHECK_CUDA(launch_kernel(finalize_token_kernel, dim3(a.B), dim3(128), 0, st, a));
  note_launch();
  return GIC_OK;
}

__global__ void init_decode_state_kernel(unsigned char* finished, int* first_eos, int B, int max_new, int* d_step, int* d_pos,
                                         int* done_counter, int P, int* fin_counter, int* all_done, int64_t* ids, int eos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    finished[i] = 0;
    first_eos[i] = max_new;
  }
  // steps that are never run (every row finished early) read as EOS, what the rows would have kept emitting (:458-460)
  for (size_t j = i; j < (size_t)B * max_new; j += (size_t)gridDim.x * blockDim.x) ids[j] = (int64_t)eos;
  if (i < 8) { fin_counter[i] = 0; all_done[i] = 0; }
  if (i == 0) {
    // counter set 0: the whole batch through prefill and token 0, then the first row group; sets 1..7: the other row groups,
    // which start decoding at step 1 with their input token at position P
    d_step[0] = 0; d_pos[0] = P - 1; done_counter[0] = 0;
    for (int s = 1; s < 8; ++s) { d_step[s] = 1; d_pos[s] = P; done_counter[s] = 0; }
  }
}

int launch_init_decode_state(unsigned char* finished, int* first_eos, int B, int max_new, int* d_step, int* d_pos, int* done_counter,
                             int P, int* fin_counter, int* all_done, int64_t* ids, int eos, cudaStream_t st) {
  init_decode_state_kernel<<<ceil_div(B, 256), 256, 0, st>>>(finished, first_eos, B, max_new, d_step, d_pos, done_counter, P, fin_counter, all_done, ids,
                                                             eos);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Compaction of finished rows: the reference finishes rows individually (src/models.py:453-460) and trained
// models stop at 10-20 of max_length = 50 tokens, so a batch soon carries mostly EOS-emitting rows.  Between two chunks of decode steps
// the live rows' per-step state (next-step input h, its bf16 / hi + lo copy, its LayerNorm statistics) is packed to the front; the KV
// cache does NOT move: slot i keeps attending cache row row_map[i].  compact_plan_kernel (one block) builds the new map with a stable
// scan over the old slots, compact_move_kernel gathers the state through a scratch copy (slots only move down, but blocks run in
// any order).  Slots [live, m_new) are padding (-1): their GEMM rows compute garbage that nothing reads.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_plan_kernel(const unsigned char* __restrict__ finished, const int* row_map_old, int m_old, int m_new,
                                                            int* row_map_new, int* src_slot /* [m_new]: old slot of new slot i, -1 = padding */) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  pdl_launch_dependents();
  pdl_wait();
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < m_old; c0 += 1024) {
    const int i = c0 + t;
    int orig = -1;
    if (i < m_old) orig = row_map_old ? row_map_old[i] : i;
    const int live = (orig >= 0 && !finished[orig]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (live) {
      const int dst = before + within;  // (dst < m_new: the host sized m_new from a live count that can only have fallen since)
      if (dst < m_new) { row_map_new[dst] = orig; src_slot[dst] = i; }
    }
    __syncthreads();
    if (t == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += s_warp[w]; s_base += tot; }
    __syncthreads();
  }
  for (int i = s_base + t; i < m_new; i += 1024) { row_map_new[i] = -1; src_slot[i] = -1; }
}

// phase 0: scratch[i] = state[src_slot[i]]; phase 1: state[i] = scratch[i]   (one block per new slot)
__global__ void __launch_bounds__(128) compact_move_kernel(const int* __restrict__ src_slot, int phase, int d, float* h, float* h_tmp, bf16* a_hi, bf16* a_hi_tmp,
                                                           bf16* a_lo, bf16* a_lo_tmp, float2* stats, float2* stats_tmp) {
  const int i = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int src = phase == 0 ? src_slot[i] : i;
  if (src_slot[i] < 0) return;
  float* hd = phase == 0 ? h_tmp : h; const float* hs = phase == 0 ? h : h_tmp;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) *reinterpret_cast<float4*>(hd + (size_t)i * d + c) = *reinterpret_cast<const float4*>(hs + (size_t)src * d + c);
  if (a_hi) {
    bf16* ad = phase == 0 ? a_hi_tmp : a_hi; const bf16* as = phase == 0 ? a_hi : a_hi_tmp;
    for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) *reinterpret_cast<uint4*>(ad + (size_t)i * d + c) = *reinterpret_cast<const uint4*>(as + (size_t)src * d + c);
  }
  if (a_lo) {
    bf16* ad = phase == 0 ? a_lo_tmp : a_lo; const bf16* as = phase == 0 ? a_lo : a_lo_tmp;
    for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) *reinterpret_cast<uint4*>(ad + (size_t)i * d + c) = *reinterpret_cast<const uint4*>(as + (size_t)src * d + c);
  }
  if (stats && threadIdx.x == 0) {
    if (phase == 0) stats_tmp[i] = stats[src]; else stats[i] = stats_tmp[i];
  }
}

int launch_compact_rows(const CompactArgs& c, cudaStream_t st) {
  GIC_REQUIRE(c.m_new > 0 && c.m_new <= c.m_old && c.d % 8 == 0, "compact_rows: bad sizes m_old=%d m_new=%d d=%d", c.m_old, c.m_new, c.d);
  GIC_CHECK_CUDA(launch_kernel(compact_plan_kernel, dim3(1), dim3(1024), 0, st, c.finished, c.row_map_old, c.m_old, c.m_new, c.row_map_new, c.src_slot));
  note_launch();
  for (int phase = 0; phase < 2; ++phase) {
    GIC_CHECK_CUDA(launch_kernel(compact_move_kernel, dim3(c.m_new), dim3(128), 0, st, (const int*)c.src_slot, phase, c.d, c.h, c.h_tmp, c.a_hi, c.a_hi_tmp, c.a_lo,
                                 c.a_lo_tmp, c.stats, c.stats_tmp));
    note_launch();
  }
  return GIC_OK;
}

// Profiling aid: occupies the stream for `cycles` SM clocks so the host can queue the whole step behind it; the queued
// kernels then run back to back and the CUDA events between them measure device time, not host launch latency.
__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}
int launch_spin(long long cycles, cudaStream_t st) {
  spin_kernel<<<1, 1, 0, st>>>(cycles);
  GIC_CHECK_CUDA(cudaGetLastError());
  return GIC_OK;
}

// L_gen of src/models.py:389-391: the reference loop stops BEFORE a step once every row has emitted EOS, so
// L_gen = max_new if some row never finished, else max_b(first_eos[b]) + 1.
__global__ void gen_len_kernel(const int* __restrict__ first_eos, int B, int max_new, int* gen_len_out) {
  __shared__ int smax[32];
  int m = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) m = max(m, first_eos[i] >= max_new ? max_new : first_eos[i] + 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = max(m, smax[w]);
    *gen_len_out = min(m, max_new);
  }
}

int launch_gen_len(const int* first_eos, int B, int max_new, int* gen_len_out, cudaStream_t st) {
  gen_len_kernel<<<1, 256, 0, st>>>(first_eos, B, max_new, gen_len_out);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

}  // namespace gic
