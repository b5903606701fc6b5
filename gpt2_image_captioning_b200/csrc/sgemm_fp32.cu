// CUDA-core fp32 GEMM for the token-exact parity mode (GIC_DTYPE_F32).
//   C[M,N] = epilogue(A[M,K] . W[N,K]^T + bias[N])
// Replaces (fp32 arithmetic of) HF Conv1D addmm (HF:pytorch_utils.py:119-123), nn.Linear and the tied LM head
// `h @ wte^T` (HF:models/gpt2/modeling_gpt2.py:705-706).  Plain FFMA with fp32 accumulation along K in order, so
// logits differ from MKL/cuBLAS only by summation order (~1e-6 relative).
// HBM-bound at decode batch sizes (weights are streamed once per step); register-tiled 8x8 / 4x4 micro-tiles,
// BK = 16, float4 global loads, double-buffered through registers.
#include "kernels.cuh"

namespace gic {

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_nt_kernel(const float* A, int lda, const float* __restrict__ W, const float* __restrict__ bias,
                float* C, int ldc, int M, int N, int K, int epilogue) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_LD = (BM * BK / 4) / NT;  // float4 loads per thread for the A tile
  constexpr int B_LD = (BN * BK / 4) / NT;
  static_assert(A_LD >= 1 && B_LD >= 1, "tile too small for the thread count");
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);
  pdl_launch_dependents();
  pdl_wait();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_LD], rb[B_LD];
  auto gload = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      int i = tid + l * NT;
      int r = i / 4, kq = (i % 4) * 4;
      int gr = m0 + r, gk = k0 + kq;
      // A is an activation produced by the previous kernel: coherent (non-invariant) load, must not be hoisted above pdl_wait
      ra[l] = (gr < M && gk < K) ? __ldcg(reinterpret_cast<const float4*>(A + (size_t)gr * lda + gk)) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int l = 0; l < B_LD; ++l) {
      int i = tid + l * NT;
      int r = i / 4, kq = (i % 4) * 4;
      int gr = n0 + r, gk = k0 + kq;
      rb[l] = (gr < N && gk < K) ? *reinterpret_cast<const float4*>(W + (size_t)gr * K + gk) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      int i = tid + l * NT;
      int r = i / 4, kq = (i % 4) * 4;
      As[buf][kq + 0][r] = ra[l].x; As[buf][kq + 1][r] = ra[l].y; As[buf][kq + 2][r] = ra[l].z; As[buf][kq + 3][r] = ra[l].w;
    }
#pragma unroll
    for (int l = 0; l < B_LD; ++l) {
      int i = tid + l * NT;
      int r = i / 4, kq = (i % 4) * 4;
      Bs[buf][kq + 0][r] = rb[l].x; Bs[buf][kq + 1][r] = rb[l].y; Bs[buf][kq + 2][r] = rb[l].z; Bs[buf][kq + 3][r] = rb[l].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + ty * TM + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx * TN + j;
      if (col >= N) continue;
      float v = acc[i][j] + (bias ? bias[col] : 0.f);
      if (epilogue == EPI_TANH) v = tanhf(v);
      else if (epilogue == EPI_GELU) v = gelu_tanh(v);
      else if (epilogue == EPI_RELU) v = fmaxf(v, 0.f);
      float* dst = C + (size_t)row * ldc + col;
      if (epilogue == EPI_RESIDUAL) v += *dst;
      *dst = v;
    }
  }
}

int launch_sgemm_nt(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N, int K, int epilogue,
                    cudaStream_t st) {
  GIC_REQUIRE(M > 0 && N > 0 && K > 0, "sgemm: empty problem M=%d N=%d K=%d", M, N, K);
  GIC_REQUIRE(K % 4 == 0 && lda % 4 == 0, "sgemm: K (%d) and lda (%d) must be multiples of 4", K, lda);
  GIC_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "sgemm: A/W must be 16-byte aligned");
  // small-M (decode at B<=64) or few tiles: 64x64 tiles put more CTAs on the 148 SMs
  const long tiles128 = (long)ceil_div(M, 128) * ceil_div(N, 128);
  if (M <= 64 || tiles128 < 2 * 148) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    GIC_CHECK_CUDA(launch_kernel(sgemm_nt_kernel<64, 64, 4, 4>, grid, dim3(256), 0, st, A, lda, W, bias, C, ldc, M, N, K, epilogue));
  } else {
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128));
    GIC_CHECK_CUDA(launch_kernel(sgemm_nt_kernel<128, 128, 8, 8>, grid, dim3(256), 0, st, A, lda, W, bias, C, ldc, M, N, K, epilogue));
  }
  note_launch();
  return GIC_OK;
}

}  // namespace gic
