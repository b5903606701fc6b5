// Row-wise / elementwise kernels of the caption path: LayerNorm, weight packing, operand staging, prefix embedding.
// All are HBM/L2-bound streaming kernels: one warp per row (LayerNorm) or one 16-byte vector per thread.
#include "kernels.cuh"

namespace gic {


// ---------------------------------------------------------------------------------------------------------------
// LayerNorm (nn.LayerNorm, eps 1e-5, biased variance): ln_1 / ln_2 / ln_f of HF GPT2Block
// (HF:models/gpt2/modeling_gpt2.py:273,304,628) and norm1/norm2 of nn.TransformerEncoderLayer.
// One warp per row, the row held in registers (d <= 32*MAXV), two-pass mean / variance in fp32.
// ---------------------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(128) layernorm_kernel(const float* x, long x_row_stride, const float* __restrict__ w,
                                                        const float* __restrict__ b, ActOut y, int rows, int d, StepTrace step_trace) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp >= rows) return;
  pdl_wait();
  const int tslot = trace_begin(step_trace, TRACE_LAYERNORM, 0);
  const float* xr = x + (size_t)warp * x_row_stride;
  // every load (row, gamma, beta) is issued before the first store: the output pointers may alias as far as the compiler
  // knows, and interleaving loads with stores serialises one memory round trip per element (ncu: 25 k cycles per warp)
  float v[MAXV], wv[MAXV], bv[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const bool ok = c < d;
    v[i] = ok ? __ldcg(xr + c) : 0.f;  // activation written by the previous kernel: NOT an invariant load (must stay below pdl_wait)
    wv[i] = ok ? __ldg(w + c) : 0.f;
    bv[i] = ok ? __ldg(b + c) : 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) s += v[i];
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float t = (c < d) ? (v[i] - mean) : 0.f;
    q += t * t;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) v[i] = (v[i] - mean) * rstd * wv[i] + bv[i];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < d) y.write((size_t)warp * d + c, v[i]);
  }
  trace_end(step_trace, tslot);
}

int launch_layernorm(const float* x, long x_row_stride, const float* w, const float* b, ActOut y, int rows, int d, cudaStream_t st) {
  GIC_REQUIRE(rows > 0 && d > 0 && d <= 32 * 40, "layernorm: unsupported rows=%d d=%d (d <= 1280)", rows, d);
  const int blocks = ceil_div(rows, 4);
  auto kern = d <= 32 * 4 ? layernorm_kernel<4> : d <= 32 * 24 ? layernorm_kernel<24> : d <= 32 * 32 ? layernorm_kernel<32> : layernorm_kernel<40>;
  GIC_CHECK_CUDA(launch_kernel(kern, dim3(blocks), dim3(128), 0, st, x, x_row_stride, w, b, y, rows, d, trace_desc()));
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight packing (engine load time): fp32 [R,C] -> optional transpose -> fp32 / bf16 hi / bf16 lo.
// HF Conv1D weights are [in,out]; every GEMM here wants W as [N,K] K-major, so Conv1D weights are transposed once.
// 32x32 smem tile transpose, coalesced on both sides.
// ---------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ in, int R, int C, bool transpose, ActOut out, const float* __restrict__ scale_k) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    // k (the in-feature index the LayerNorm gamma scales) is the source ROW when the weight is transposed, else the column
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * C + c] * (scale_k ? scale_k[transpose ? r : c] : 1.0f) : 0.f;
  }
  __syncthreads();
  if (!transpose) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int r = r0 + i, c = c0 + threadIdx.x;
      if (r < R && c < C) out.write((size_t)r * C + c, tile[i][threadIdx.x]);
    }
  } else {  // out is [C, R]
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < R && c < C) out.write((size_t)c * R + r, tile[threadIdx.x][i]);
    }
  }
}

int launch_pack_weight(const float* in, int R, int C, bool transpose, ActOut out, cudaStream_t st, const float* scale_k) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
  GIC_REQUIRE(grid.y <= 65535, "pack_weight: too many rows (%d)", R);
  pack_weight_kernel<<<grid, block, 0, st>>>(in, R, C, transpose, out, scale_k);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm folded into the following GEMM (bf16 engine, decode + prefill):
//   LN(x) . W^T + b = rstd * (x . (gamma * W)^T - mean * colsum) + (b + W . beta),   colsum[n] = sum_k gamma_k W[n,k]
// The GEMM consumes the RAW rows (bf16 copy of the residual stream) and applies mean / rstd per row in its epilogue, so the
// 25 LayerNorm launches of a decode step disappear (HF GPT2Block ln_1 / ln_2 and ln_f, HF:models/gpt2/modeling_gpt2.py:273,304,628).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fold_ln_kernel(const bf16* __restrict__ w_packed, const bf16* __restrict__ w_packed_lo, const float* __restrict__ w_src,
                                                      bool transposed, const float* __restrict__ beta, const float* __restrict__ bias,
                                                      float* __restrict__ colsum, float* __restrict__ bias_out, int N, int K) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    cs += __bfloat162float(w_packed[(size_t)n * K + k]);  // what the tensor core will really multiply by
    if (w_packed_lo) cs += __bfloat162float(w_packed_lo[(size_t)n * K + k]);  // (bf16x2: hi + lo)
    bs += beta[k] * (transposed ? w_src[(size_t)k * N + n] : w_src[(size_t)n * K + k]);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_out[n] = (bias ? bias[n] : 0.f) + bs;
  }
}

int launch_fold_ln(const bf16* w_packed, const float* w_src, bool transposed, const float* beta, const float* bias, float* colsum,
                   float* bias_out, int N, int K, cudaStream_t st, const bf16* w_packed_lo) {
  fold_ln_kernel<<<ceil_div(N, 4), 128, 0, st>>>(w_packed, w_packed_lo, w_src, transposed, beta, bias, colsum, bias_out, N, K);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

template <int MAXV>
__global__ void __launch_bounds__(128) row_stats_kernel(const float* x, long x_row_stride, bf16* __restrict__ xb, bf16* __restrict__ xb_lo,
                                                        float2* __restrict__ stats, int rows, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp >= rows) return;
  pdl_wait();
  const float* xr = x + (size_t)warp * x_row_stride;
  float v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    v[i] = c < d ? __ldcg(xr + c) : 0.f;
  }
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const bf16 b = __float2bfloat16_rn(v[i]);
    float f = __bfloat162float(b);
    if (c < d) {
      xb[(size_t)warp * d + c] = b;
      if (xb_lo) {  // bf16x2: the operand is hi + lo (exact in fp32), and the statistics are those of the sum
        const bf16 l = __float2bfloat16_rn(v[i] - f);
        xb_lo[(size_t)warp * d + c] = l;
        f += __bfloat162float(l);
      }
      s += f;
      q += f * f;
    }
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) stats[warp] = make_float2(s, q);
}

int launch_row_stats(const float* x, long x_row_stride, bf16* xb, float2* stats, int rows, int d, cudaStream_t st, bf16* xb_lo) {
  GIC_REQUIRE(rows > 0 && d > 0 && d <= 32 * 40, "row_stats: unsupported rows=%d d=%d (d <= 1280)", rows, d);
  const int blocks = ceil_div(rows, 4);
  auto kern = d <= 32 * 4 ? row_stats_kernel<4> : d <= 32 * 24 ? row_stats_kernel<24> : d <= 32 * 32 ? row_stats_kernel<32> : row_stats_kernel<40>;
  GIC_CHECK_CUDA(launch_kernel(kern, dim3(blocks), dim3(128), 0, st, x, x_row_stride, xb, xb_lo, stats, rows, d));
  note_launch();
  return GIC_OK;
}

__global__ void convert_kernel(const float* __restrict__ in, ActOut out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out.write(i, in[i]);
}

int launch_convert(const float* in, ActOut out, size_t n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  convert_kernel<<<blocks, 256, 0, st>>>(in, out, n);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// h[b,t,:] = prefix_token[b,t,:] + wpe[t,:]  -- HF GPT2Model.forward :579-585 (prefix tokens also get positions);
// prefix tokens = image prefix [B,P_img,d] followed by the task prefix [P_task,d] (src/models.py:364-375).
// ---------------------------------------------------------------------------------------------------------------
__global__ void embed_prefix_kernel(const float* __restrict__ prefix, int P_img, const float* __restrict__ task, int P_task,
                                    const float* __restrict__ wpe, float* __restrict__ h, float* __restrict__ prefix_out, int B,
                                    int d) {
  const int P = P_img + P_task;
  const size_t n4 = (size_t)B * P * d / 4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const size_t e = i * 4;
    const int c = (int)(e % d);
    const int t = (int)((e / d) % P);
    const int bb = (int)(e / ((size_t)d * P));
    float4 v = (t < P_img) ? *reinterpret_cast<const float4*>(prefix + ((size_t)bb * P_img + t) * d + c)
                           : *reinterpret_cast<const float4*>(task + (size_t)(t - P_img) * d + c);
    if (prefix_out) *reinterpret_cast<float4*>(prefix_out + e) = v;
    if (h) {
      const float4 p = *reinterpret_cast<const float4*>(wpe + (size_t)t * d + c);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      *reinterpret_cast<float4*>(h + e) = v;
    }
  }
}

int launch_embed_prefix(const float* prefix, int P_img, const float* task, int P_task, const float* wpe, float* h, float* prefix_out,
                        int B, int d, cudaStream_t st) {
  GIC_REQUIRE(d % 4 == 0, "embed_prefix: d (%d) must be a multiple of 4", d);
  const size_t n4 = (size_t)B * (P_img + P_task) * d / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_prefix_kernel<<<blocks, 256, 0, st>>>(prefix, P_img, task, P_task, wpe, h, prefix_out, B, d);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// rows [B, S_src, d] (fp32) -> gather a token window [B, n_tok, d] starting at tok0 (transformer mapper: last P tokens)
__global__ void slice_tokens_kernel(const float* __restrict__ in, int S, int tok0, int n_tok, float* __restrict__ out, int B, int d) {
  const size_t n = (size_t)B * n_tok * d;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int c = (int)(i % d);
    const int t = (int)((i / d) % n_tok);
    const int bb = (int)(i / ((size_t)d * n_tok));
    out[i] = in[((size_t)bb * S + tok0 + t) * d + c];
  }
}

int launch_slice_tokens(const float* in, int S, int tok0, int n_tok, float* out, int B, int d, cudaStream_t st) {
  const size_t n = (size_t)B * n_tok * d;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  slice_tokens_kernel<<<blocks, 256, 0, st>>>(in, S, tok0, n_tok, out, B, d);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

// transformer mapper input sequence: seq[b, t, :] = (t < Hl) ? lin[b, t*d .. ] : prefix_const[t-Hl]   (src/models.py:157-168)
__global__ void build_mapper_seq_kernel(const float* __restrict__ lin, const float* __restrict__ prefix_const, float* __restrict__ seq,
                                        int B, int Hl, int P, int d) {
  const int S = Hl + P;
  const size_t n = (size_t)B * S * d;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int c = (int)(i % d);
    const int t = (int)((i / d) % S);
    const int bb = (int)(i / ((size_t)d * S));
    seq[i] = (t < Hl) ? lin[((size_t)bb * Hl + t) * d + c] : prefix_const[(size_t)(t - Hl) * d + c];
  }
}

int launch_build_mapper_seq(const float* lin, const float* prefix_const, float* seq, int B, int Hl, int P, int d, cudaStream_t st) {
  const size_t n = (size_t)B * (Hl + P) * d;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  build_mapper_seq_kernel<<<blocks, 256, 0, st>>>(lin, prefix_const, seq, B, Hl, P, d);
  GIC_CHECK_CUDA(cudaGetLastError());
  note_launch();
  return GIC_OK;
}

}  // namespace gic
