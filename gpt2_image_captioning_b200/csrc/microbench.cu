// Stand-alone kernel microbenchmarks (not part of the library): `make microbench && build/microbench`.
// Times the decode-step kernels of the headline workload in isolation with CUDA events, rotating over 12 weight /
// cache sets so the working set exceeds L2 like a real step, plus a few hardware probes (launch floor, SM clock,
// dependent-load latency).
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "kernels.cuh"

using namespace gic;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)
#define OK(x)                                                              \
  do {                                                                     \
    if ((x) != GIC_OK) { printf("gic error: %s (%s:%d)\n", get_error(), __FILE__, __LINE__); exit(1); } \
  } while (0)

namespace gic {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
const char* get_error() { return g_err; }
void note_launch() {}
StepTrace trace_desc() { return StepTrace{nullptr, 0, 0}; }
bool trace_on() { return false; }
unsigned int trace_generation() { return 0; }
static int g_cta_limit = 0;
int cta_limit() { return g_cta_limit; }
void set_cta_limit(int c) { g_cta_limit = c; }
bool pdl_enabled() { static int on = -1; if (on < 0) { const char* v = getenv("GIC_NO_PDL"); on = (v && v[0] == '1') ? 0 : 1; } return on == 1; }
}  // namespace gic

__global__ void empty_kernel() {}
__global__ void clock_probe(long long* out) {
  long long c0 = clock64();
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 200000);  // 200 us
  out[0] = clock64() - c0;
  out[1] = (long long)(t1 - t0);
}
__global__ void chase_kernel(const int* __restrict__ next, int start, int n, long long* out) {
  int p = start;
  long long c0 = clock64();
  for (int i = 0; i < n; ++i) p = next[p];
  out[0] = clock64() - c0;
  out[1] = p;
}

__device__ __forceinline__ float hash_unit(size_t i, unsigned seed) {  // (-1, 1)
  unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
  x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
  return (float)((int)(x & 0xffffff) - 0x800000) * (1.0f / 0x800000);
}
__global__ void fill_bf16(bf16* p, size_t n, unsigned seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = __float2bfloat16_rn(scale * hash_unit(i, seed));
}
__global__ void fill_f32(float* p, size_t n, unsigned seed, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = scale * hash_unit(i, seed);
}

template <typename F>
static float time_loop(cudaStream_t st, int iters, F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f(i);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(a, st));
  for (int i = 0; i < iters; ++i) f(i);
  CK(cudaEventRecord(b, st));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / iters;  // us per iteration
}

__global__ void stream_read_kernel(const uint4* p, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i < n16; i += stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u * blockDim.x < n16) ? __ldcs(p + i + u * blockDim.x) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024;
  const int d = 768, H = 12, L = 12, V = 50257;
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clockRate %d kHz, memClock %d kHz, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.clockRate,
         prop.memoryClockRate, prop.l2CacheSize >> 20);

  // ---- probes ----
  printf("empty kernel back-to-back: %.2f us/launch\n", time_loop(st, 2000, [&](int) { empty_kernel<<<1, 32, 0, st>>>(); }));
  long long* dout; CK(cudaMalloc(&dout, 16));
  clock_probe<<<1, 1, 0, st>>>(dout);
  long long hout[2]; CK(cudaMemcpyAsync(hout, dout, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  printf("SM clock under a 1-thread spin: %.0f MHz\n", (double)hout[0] / ((double)hout[1] / 1000.0));
  {
    const int n = 1 << 24;  // 64 MB of ints: mostly L2-resident after warm-up? (126 MB L2) -> use stride to defeat
    std::vector<int> h(n);
    const int stride = 4099 * 32;
    for (int i = 0; i < n; ++i) h[i] = (int)(((long long)i + stride) % n);
    int* dn; CK(cudaMalloc(&dn, (size_t)n * 4));
    CK(cudaMemcpy(dn, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    chase_kernel<<<1, 1, 0, st>>>(dn, 0, 2000, dout);
    CK(cudaMemcpyAsync(hout, dout, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    printf("dependent global load latency (first touch): %.0f cycles\n", (double)hout[0] / 2000);
    chase_kernel<<<1, 1, 0, st>>>(dn, 0, 2000, dout);
    CK(cudaMemcpyAsync(hout, dout, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    printf("dependent global load latency (second pass, L2): %.0f cycles\n", (double)hout[0] / 2000);
    cudaFree(dn);
  }

  // ---- buffers: 12 rotating sets ----
  const int SETS = 12;
  auto dmalloc = [&](size_t bytes) { void* p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes)); return p; };
  float* h = (float*)dmalloc((size_t)B * d * 4);
  float* lnw = (float*)dmalloc(d * 4); float* lnb = (float*)dmalloc(d * 4);
  bf16* a = (bf16*)dmalloc((size_t)B * 4 * d * 2);
  bf16* qkv = (bf16*)dmalloc((size_t)B * 3 * d * 2);
  bf16* o = (bf16*)dmalloc((size_t)B * 4 * d * 2);
  float* bias = (float*)dmalloc(4 * d * 4);
  int* dpos = (int*)dmalloc(4);
  const int t_max = 40;
  std::vector<bf16*> w_qkv(SETS), w_proj(SETS), w_fc(SETS), w_fc2(SETS), kc(SETS), vc(SETS);
  for (int s = 0; s < SETS; ++s) {
    w_qkv[s] = (bf16*)dmalloc((size_t)3 * d * d * 2); w_proj[s] = (bf16*)dmalloc((size_t)d * d * 2);
    w_fc[s] = (bf16*)dmalloc((size_t)4 * d * d * 2); w_fc2[s] = (bf16*)dmalloc((size_t)4 * d * d * 2);
    kc[s] = (bf16*)dmalloc((size_t)B * H * t_max * 64 * 2); vc[s] = (bf16*)dmalloc((size_t)B * H * t_max * 64 * 2);
  }
  bf16* wte = (bf16*)dmalloc((size_t)V * d * 2);
  float* pv = (float*)dmalloc((size_t)2048 * B * 4); int* pi = (int*)dmalloc((size_t)2048 * B * 4);
  OK(tma_init()); OK(gemm_bf16_configure()); OK(attn_decode_configure());

  // ---- mode 5 <variant> <ctx>: a handful of decode-attention launches of one kernel variant (the ncu target) ----
  if (argc > 4 && atoi(argv[2]) == 5) {
    const int variant = atoi(argv[3]), ctx = atoi(argv[4]);
    int pos = ctx - 1;
    CK(cudaMemcpy(dpos, &pos, 4, cudaMemcpyHostToDevice));
    attn_decode_set_variant(variant);
    ActOut y; y.hi = o;
    float us = time_loop(st, 6, [&](int i) { OK(launch_attn_decode<bf16>(qkv, kc[i % SETS], vc[i % SETS], y, dpos, B, H, t_max, st)); });
    printf("attn_decode v%d rows=%d ctx=%d: %.2f us (6 launches)\n", variant, B, ctx, us);
    return 0;
  }

  // ---- tensor-pipe probe: the LM-head GEMM with every MMA re-issued R times (no extra operand traffic) ----
  if (argc > 2 && atoi(argv[2]) == 2) {
    for (int bn : {64, 128, 256}) {
      float base = 0;
      for (int rep : {1, 3, 5}) {
        GemmBf16Args g;
        OK(make_tma_2d_bf16(&g.a_hi, a, B, d, d, 128));
        OK(make_tma_2d_bf16(&g.w_hi, wte, V, d, d, bn));
        g.M = B; g.N = V; g.K = d; g.block_n = bn; g.part_val = pv; g.part_idx = pi; g.part_ld = 2048; g.mma_repeat = rep;
        float us = time_loop(st, 20, [&](int) { OK(launch_gemm_bf16(g, st)); });
        if (rep == 1) base = us;
        const double tiles_per_cta = (double)((B + 127) / 128) * ((V + bn - 1) / bn) / 148.0;
        const double mmas_per_cta = tiles_per_cta * (d / 16);
        printf("mma_probe block_n=%3d repeat=%d: %8.2f us", bn, rep, us);
        if (rep > 1) printf("   -> %.0f cycles per extra 128x%dx16 MMA (floor %d)", (us - base) * 1965.0 / ((rep - 1) * mmas_per_cta), bn, bn / 2);
        printf("\n");
      }
    }
    return 0;
  }

  // ---- epilogue elimination: LM-head shape with (a) no outputs at all, (b) argmax partials, (c) bf16 logits stored ----
  if (argc > 2 && atoi(argv[2]) == 3) {
    bf16* big = (bf16*)dmalloc((size_t)B * V * 2 + 1024);
    for (int bn : {128, 256}) for (int mode = 0; mode < 3; ++mode) {
      GemmBf16Args g;
      OK(make_tma_2d_bf16(&g.a_hi, a, B, d, d, 128));
      OK(make_tma_2d_bf16(&g.w_hi, wte, V, d, d, bn));
      g.M = B; g.N = V; g.K = d; g.block_n = bn;
      if (mode == 1) { g.part_val = pv; g.part_idx = pi; g.part_ld = 2048; }
      if (mode == 2) { g.out.hi = big; g.ld_out = V; }
      float us = time_loop(st, 20, [&](int) { OK(launch_gemm_bf16(g, st)); });
      printf("epi_probe block_n=%3d %-14s: %8.2f us  %.1f TFLOP/s\n", bn, mode == 0 ? "no output" : mode == 1 ? "argmax" : "bf16 stores", us,
             2.0 * B * V * d / us * 1e-6);
    }
    return 0;
  }

  // ---- main-loop timeline of CTA 0 (clock64 deltas from kernel entry) ----
  if (argc > 2 && atoi(argv[2]) == 4) {
    long long* tr = (long long*)dmalloc(640 * 8);
    struct Shape { const char* name; bf16* A; bf16* W; int N, K, bn; bool lm; int sk; };
    float* skws = (float*)dmalloc((size_t)4 * B * d * 4);
    int* skcnt = (int*)dmalloc(4096 * 4);
    float2* stats = (float2*)dmalloc((size_t)64 * B * 8);
    Shape shapes[] = {{"qkv", a, w_qkv[0], 3 * d, d, 128, false, 1}, {"fc2", a, w_fc2[0], d, 4 * d, 64, false, 1}, {"fc2 split-K 3", a, w_fc2[0], d, 4 * d, 128, false, 3},
                      {"lm_head", a, wte, V, d, 128, true, 1}};
    for (const Shape& sh : shapes) {
      GemmBf16Args g;
      OK(make_tma_2d_bf16(&g.a_hi, sh.A, B, sh.K, sh.K, 128));
      OK(make_tma_2d_bf16(&g.w_hi, sh.W, sh.N, sh.K, sh.K, sh.bn));
      g.M = B; g.N = sh.N; g.K = sh.K; g.block_n = sh.bn; g.bias = bias;
      if (sh.lm) { g.part_val = pv; g.part_idx = pi; g.part_ld = 2048; g.bias = nullptr; } else { g.out.hi = o; g.ld_out = sh.N; }
      if (sh.sk > 1) { g.epilogue = EPI_RESIDUAL; g.out.f32 = h; g.stats_out = stats; g.ln_stats_ld = B; g.split_k = sh.sk; g.splitk_ws = skws; g.splitk_counters = skcnt; }
      for (int i = 0; i < 3; ++i) OK(launch_gemm_bf16(g, st));
      CK(cudaMemsetAsync(tr, 0, 640 * 8, st));
      g.trace = tr;
      OK(launch_gemm_bf16(g, st));
      long long h_tr[640];
      CK(cudaMemcpyAsync(h_tr, tr, sizeof(h_tr), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      printf("timeline %s (N=%d K=%d block_n=%d): prologue done at %lld cycles\n", sh.name, sh.N, sh.K, sh.bn, h_tr[600]);
      printf("  kb | producer: wait_begin empty_seen issued | mma: wait_begin full_seen committed\n");
      const int nkb = sh.K / 64 * (sh.lm ? 3 : 1);
      for (int kb = 0; kb < nkb && kb < 60; ++kb)
        printf("  %2d | %8lld %8lld %8lld | %8lld %8lld %8lld\n", kb, h_tr[kb * 4], h_tr[kb * 4 + 1], h_tr[kb * 4 + 2], h_tr[256 + kb * 4],
               h_tr[256 + kb * 4 + 1], h_tr[256 + kb * 4 + 2]);
      for (int t = 0; t < (sh.lm ? 3 : 1); ++t)
        printf("  epilogue tile %d: wait_begin %lld  accumulator_seen %lld  done %lld\n", t, h_tr[512 + t * 4], h_tr[512 + t * 4 + 1], h_tr[512 + t * 4 + 2]);
      for (int c = 0; c < 2; ++c)
        printf("  epilogue warp 0 chunk %d: begin %lld  tmem_loaded %lld  staged %lld  bias_loaded %lld  stored %lld\n", c, h_tr[540 + c * 8], h_tr[541 + c * 8],
               h_tr[542 + c * 8], h_tr[543 + c * 8], h_tr[544 + c * 8]);
    }
    return 0;
  }

  // ---- mode 6: the bf16x2 (hi + lo operands, 3 MMAs per product) decode GEMMs in their fused form over every tile shape, the split
  // LM head, and the fp16-cache decode attention next to the bf16 one ----
  if (argc > 2 && atoi(argv[2]) == 6) {
    bf16* a_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    bf16* o_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    std::vector<bf16*> l_qkv(SETS), l_proj(SETS), l_fc(SETS), l_fc2(SETS);
    for (int s = 0; s < SETS; ++s) {
      l_qkv[s] = (bf16*)dmalloc((size_t)3 * d * d * 2); l_proj[s] = (bf16*)dmalloc((size_t)d * d * 2);
      l_fc[s] = (bf16*)dmalloc((size_t)4 * d * d * 2); l_fc2[s] = (bf16*)dmalloc((size_t)4 * d * d * 2);
    }
    bf16* wte_lo = (bf16*)dmalloc((size_t)V * d * 2);
    float2* stats = (float2*)dmalloc((size_t)64 * B * 8);
    float* colsum = (float*)dmalloc((size_t)V * 4);
    struct Shape { const char* name; int N, K; std::vector<bf16*>*w, *wl; int epi; bool res; };
    Shape shapes[] = {{"qkv", 3 * d, d, &w_qkv, &l_qkv, EPI_NONE, false}, {"proj", d, d, &w_proj, &l_proj, EPI_RESIDUAL, true},
                      {"fc", 4 * d, d, &w_fc, &l_fc, EPI_GELU, false}, {"fc2", d, 4 * d, &w_fc2, &l_fc2, EPI_RESIDUAL, true}};
    for (auto& s : shapes) {
      int bn_pick = 0, pair_pick = 0;
      gemm_bf16_pick(B, s.N, s.K, 1, 1, &bn_pick, &pair_pick);
      for (int pair = 0; pair < 2; ++pair)
        for (int bn : {64, 128, 192, 256}) {
          if (!pair && bn > 128) continue;
          if (pair && ((B + 127) / 128) % 2) continue;
          std::vector<GemmBf16Args> args(SETS);
          for (int i = 0; i < SETS; ++i) {
            GemmBf16Args& g = args[i];
            OK(make_tma_2d_bf16(&g.a_hi, a, B, s.K, s.K, 128));
            OK(make_tma_2d_bf16(&g.a_lo, a_lo, B, s.K, s.K, 128));
            OK(make_tma_2d_bf16(&g.w_hi, (*s.w)[i], s.N, s.K, s.K, pair ? bn / 2 : bn));
            OK(make_tma_2d_bf16(&g.w_lo, (*s.wl)[i], s.N, s.K, s.K, pair ? bn / 2 : bn));
            g.M = B; g.N = s.N; g.K = s.K; g.block_n = bn; g.split = 1; g.epilogue = s.epi; g.bias = bias; g.ld_out = s.N; g.pair = pair; g.w_static = 1;
            if (s.res) { g.out.f32 = h; g.out.hi = o; g.out.lo = o_lo; g.stats_out = stats; g.ln_stats_ld = B; }
            else {
              g.ln_stats = stats; g.ln_parts = d / 32; g.ln_stats_ld = B; g.ln_colsum = colsum;
              if (s.epi == EPI_NONE) { g.out.hi = qkv; g.out_f16 = 1; } else { g.out.hi = o; g.out.lo = o_lo; }
            }
          }
          float us = time_loop(st, 240, [&](int i) { OK(launch_gemm_bf16(args[i % SETS], st)); });
          printf("x2 gemm %-4s%s M=%d N=%d K=%d block_n=%3d: %7.2f us  %6.1f TFLOP/s algorithmic (x3 on the pipe)%s\n", s.name, pair ? " pair" : "     ", B, s.N, s.K, bn, us,
                 2.0 * B * s.N * s.K / us * 1e-6, (bn == bn_pick && pair == pair_pick) ? "  <- picked" : "");
        }
    }
    for (int pair = 0; pair < 2; ++pair)
      for (int bn : {128, 192, 256}) {
        if (!pair && bn > 128) continue;
        GemmBf16Args g;
        OK(make_tma_2d_bf16(&g.a_hi, a, B, d, d, 128));
        OK(make_tma_2d_bf16(&g.a_lo, a_lo, B, d, d, 128));
        OK(make_tma_2d_bf16(&g.w_hi, wte, V, d, d, pair ? bn / 2 : bn));
        OK(make_tma_2d_bf16(&g.w_lo, wte_lo, V, d, d, pair ? bn / 2 : bn));
        g.M = B; g.N = V; g.K = d; g.block_n = bn; g.split = 1; g.part_val = pv; g.part_idx = pi; g.part_ld = 2048; g.pair = pair;
        float us = time_loop(st, 30, [&](int) { OK(launch_gemm_bf16(g, st)); });
        printf("x2 lm_head%s M=%d N=%d K=%d block_n=%d: %.2f us  %.1f TFLOP/s algorithmic\n", pair ? " pair" : "     ", B, V, d, bn, us, 2.0 * B * V * d / us * 1e-6);
      }
    for (int f16 = 0; f16 < 2; ++f16)
      for (int ctx : {11, 25, 39}) {
        int pos = ctx - 1;
        CK(cudaMemcpy(dpos, &pos, 4, cudaMemcpyHostToDevice));
        ActOut y; y.hi = o;
        float us = time_loop(st, 240, [&](int i) {
          if (f16) OK(launch_attn_decode_f16(qkv, kc[i % SETS], vc[i % SETS], o, o_lo, dpos, B, H, t_max, st));
          else OK(launch_attn_decode<bf16>(qkv, kc[i % SETS], vc[i % SETS], y, dpos, B, H, t_max, st));
        });
        const double bytes = 2.0 * B * d * (2.0 * ctx + 2 + 3 + 1 + (f16 ? 1 : 0));
        printf("attn_decode %s rows=%d ctx=%d: %.2f us  %.0f GB/s\n", f16 ? "fp16 cache, hi+lo out" : "bf16                 ", B, ctx, us, bytes / us * 1e-3);
      }
    return 0;
  }

  // ---- mode 7: CTA-0 clock64 timeline of the bf16x2 decode GEMMs in their engine configuration (tile / pair as the engine picks) ----
  if (argc > 2 && atoi(argv[2]) == 7) {
    long long* tr = (long long*)dmalloc(640 * 8);
    bf16* a_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    bf16* o_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    bf16* wl = (bf16*)dmalloc((size_t)4 * d * d * 2);
    float2* stats = (float2*)dmalloc((size_t)64 * B * 8);
    float* colsum = (float*)dmalloc((size_t)V * 4);
    struct Shape { const char* name; bf16* W; int N, K, epi; bool res; int sk; };
    float* skws = (float*)dmalloc((size_t)4 * B * d * 4);
    int* skcnt = (int*)dmalloc(4096 * 4);
    Shape shapes[] = {{"qkv", w_qkv[0], 3 * d, d, EPI_NONE, false, 1}, {"proj", w_proj[0], d, d, EPI_RESIDUAL, true, 1},
                      {"fc", w_fc[0], 4 * d, d, EPI_GELU, false, 1}, {"fc2", w_fc2[0], d, 4 * d, EPI_RESIDUAL, true, 1},
                      {"fc2 split-K 3", w_fc2[0], d, 4 * d, EPI_RESIDUAL, true, 3}, {"proj split-K 2", w_proj[0], d, d, EPI_RESIDUAL, true, 2}};
    for (const Shape& sh : shapes) {
      int bn = 0, pair = 0;
      gemm_bf16_pick(B, sh.N, sh.K, 1, sh.sk, &bn, sh.sk > 1 ? nullptr : &pair);
      GemmBf16Args g;
      OK(make_tma_2d_bf16(&g.a_hi, a, B, sh.K, sh.K, 128));
      OK(make_tma_2d_bf16(&g.a_lo, a_lo, B, sh.K, sh.K, 128));
      OK(make_tma_2d_bf16(&g.w_hi, sh.W, sh.N, sh.K, sh.K, pair ? bn / 2 : bn));
      OK(make_tma_2d_bf16(&g.w_lo, wl, sh.N, sh.K, sh.K, pair ? bn / 2 : bn));
      g.M = B; g.N = sh.N; g.K = sh.K; g.block_n = bn; g.split = 1; g.epilogue = sh.epi; g.bias = bias; g.ld_out = sh.N; g.pair = pair; g.w_static = 1;
      if (sh.res) { g.out.f32 = h; g.out.hi = o; g.out.lo = o_lo; g.stats_out = stats; g.ln_stats_ld = B; }
      else {
        g.ln_stats = stats; g.ln_parts = d / 32; g.ln_stats_ld = B; g.ln_colsum = colsum;
        if (sh.epi == EPI_NONE) { g.out.hi = qkv; g.out_f16 = 1; } else { g.out.hi = o; g.out.lo = o_lo; }
      }
      if (sh.sk > 1) { g.split_k = sh.sk; g.splitk_ws = skws; g.splitk_counters = skcnt; }
      for (int i = 0; i < 3; ++i) OK(launch_gemm_bf16(g, st));
      float us = time_loop(st, 50, [&](int) { OK(launch_gemm_bf16(g, st)); });
      CK(cudaMemsetAsync(tr, 0, 640 * 8, st));
      g.trace = tr;
      OK(launch_gemm_bf16(g, st));
      long long h_tr[640];
      CK(cudaMemcpyAsync(h_tr, tr, sizeof(h_tr), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      const int nkb = sh.K / 64;
      printf("x2 timeline %s (N=%d K=%d block_n=%d%s, %.2f us back to back): prologue done %lld | first TMA issued %lld | first MMA issued %lld | last MMA committed %lld | "
             "epilogue: accumulator seen %lld, done %lld cycles\n", sh.name, sh.N, sh.K, bn, pair ? " pair" : "", us, h_tr[600], h_tr[2], h_tr[256 + 1],
             h_tr[256 + (nkb - 1 < 59 ? nkb - 1 : 59) * 4 + 2], h_tr[512 + 1], h_tr[512 + 2]);
      printf("   per k-block MMA issue times:");
      for (int kb = 0; kb < nkb && kb < 16; ++kb) printf(" %lld", h_tr[256 + kb * 4 + 1]);
      printf("\n   epilogue warp 0 chunk 0: begin %lld tmem_loaded %lld staged %lld bias_arrived %lld math_done %lld stores_issued %lld chunk_done(stats) %lld\n", h_tr[540], h_tr[541],
             h_tr[542], h_tr[543], h_tr[545], h_tr[546], h_tr[544]);
    }
    return 0;
  }

  // ---- mode 8: K split of the residual GEMMs (proj, fc2) with the cooperative reduction: every (operand mode, pair, tile width, split)
  // timed back to back over rotating weight sets, and its h / hi / lo / statistics checked against the unsplit 64-wide launch ----
  if (argc > 2 && atoi(argv[2]) == 8) {
    bf16* a_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    bf16* o_lo = (bf16*)dmalloc((size_t)B * 4 * d * 2);
    bf16* wl = (bf16*)dmalloc((size_t)4 * d * d * 2);
    float2* stats = (float2*)dmalloc((size_t)64 * B * 8);
    float* skws = (float*)dmalloc((size_t)4 * B * d * 4);
    int* skcnt = (int*)dmalloc(4096 * 4);
    float* h0 = (float*)dmalloc((size_t)B * d * 4);
    fill_bf16<<<512, 256, 0, st>>>(a, (size_t)B * 4 * d, 1u, 1.0f);
    fill_bf16<<<512, 256, 0, st>>>(a_lo, (size_t)B * 4 * d, 2u, 1.0f / 256);
    fill_bf16<<<512, 256, 0, st>>>(w_proj[0], (size_t)d * d, 3u, 0.05f);
    fill_bf16<<<512, 256, 0, st>>>(w_fc2[0], (size_t)4 * d * d, 4u, 0.05f);
    fill_bf16<<<512, 256, 0, st>>>(wl, (size_t)4 * d * d, 5u, 0.05f / 256);
    fill_f32<<<512, 256, 0, st>>>(h0, (size_t)B * d, 6u, 1.0f);
    fill_f32<<<512, 256, 0, st>>>(bias, (size_t)4 * d, 7u, 0.1f);
    CK(cudaStreamSynchronize(st));
    std::vector<float> ref_h((size_t)B * d), got_h((size_t)B * d);
    std::vector<float2> ref_st((size_t)(d / 32) * B), got_st((size_t)(d / 32) * B);
    std::vector<uint16_t> ref_hi((size_t)B * d), got_hi((size_t)B * d), ref_lo((size_t)B * d), got_lo((size_t)B * d);
    struct Shape { const char* name; bf16* W; int K; };
    Shape shapes[] = {{"proj", w_proj[0], d}, {"fc2", w_fc2[0], 4 * d}};
    for (int x2 = 1; x2 >= 0; --x2)
      for (const Shape& sh : shapes) {
        bool have_ref = false;
        for (int pair = 0; pair < 2; ++pair)
          for (int bn : {64, 128, 192, 256})
            for (int sk = 1; sk <= 4; ++sk) {
              if (!pair && x2 && bn > 128) continue;
              if (pair && ((B + 127) / 128) % 2) continue;
              if (sk > 1 && bn < 128) continue;
              const long work = (long)((B + 127) / 128) * ((d + bn - 1) / bn) * sk;
              if (sk > 1 && work > 148) continue;
              GemmBf16Args g;
              OK(make_tma_2d_bf16(&g.a_hi, a, B, sh.K, sh.K, 128));
              OK(make_tma_2d_bf16(&g.a_lo, a_lo, B, sh.K, sh.K, 128));
              OK(make_tma_2d_bf16(&g.w_hi, sh.W, d, sh.K, sh.K, pair ? bn / 2 : bn));
              OK(make_tma_2d_bf16(&g.w_lo, wl, d, sh.K, sh.K, pair ? bn / 2 : bn));
              g.M = B; g.N = d; g.K = sh.K; g.block_n = bn; g.split = x2; g.epilogue = EPI_RESIDUAL; g.bias = bias; g.ld_out = d; g.pair = pair; g.w_static = 1;
              g.out.f32 = h; g.out.hi = o; g.out.lo = x2 ? o_lo : nullptr; g.stats_out = stats; g.ln_stats_ld = B;
              if (sk > 1) { g.split_k = sk; g.splitk_ws = skws; g.splitk_counters = skcnt; }
              CK(cudaMemcpyAsync(h, h0, (size_t)B * d * 4, cudaMemcpyDeviceToDevice, st));
              OK(launch_gemm_bf16(g, st));
              std::vector<float>& dst_h = have_ref ? got_h : ref_h;
              std::vector<float2>& dst_st = have_ref ? got_st : ref_st;
              std::vector<uint16_t>& dst_hi = have_ref ? got_hi : ref_hi; std::vector<uint16_t>& dst_lo = have_ref ? got_lo : ref_lo;
              CK(cudaMemcpyAsync(dst_h.data(), h, (size_t)B * d * 4, cudaMemcpyDeviceToHost, st));
              CK(cudaMemcpyAsync(dst_st.data(), stats, (size_t)(d / 32) * B * 8, cudaMemcpyDeviceToHost, st));
              CK(cudaMemcpyAsync(dst_hi.data(), o, (size_t)B * d * 2, cudaMemcpyDeviceToHost, st));
              if (x2) CK(cudaMemcpyAsync(dst_lo.data(), o_lo, (size_t)B * d * 2, cudaMemcpyDeviceToHost, st));
              CK(cudaStreamSynchronize(st));
              double err_h = 0, err_st = 0, mag = 0; long bad_hi = 0;
              if (have_ref) {
                for (size_t i = 0; i < ref_h.size(); ++i) { err_h = fmax(err_h, fabs((double)got_h[i] - ref_h[i])); mag = fmax(mag, fabs((double)ref_h[i])); }
                for (size_t i = 0; i < ref_st.size(); ++i) err_st = fmax(err_st, fabs((double)got_st[i].x - ref_st[i].x) / (1.0 + fabs((double)ref_st[i].x)));
                for (size_t i = 0; i < ref_hi.size(); ++i) bad_hi += (got_hi[i] != ref_hi[i]);
              }
              have_ref = true;
              float us = time_loop(st, 120, [&](int) { OK(launch_gemm_bf16(g, st)); });
              int cnt_bad = 0;
              std::vector<int> hc(4096);
              CK(cudaMemcpy(hc.data(), skcnt, 4096 * 4, cudaMemcpyDeviceToHost));
              for (int v : hc) cnt_bad += (v != 0);
              printf("splitk %s %-4s%s block_n=%3d split_k=%d (%3ld CTAs): %7.2f us   max|dh| %.2e of %.1f  stats rel %.1e  hi differs %ld  counters nonzero %d\n", x2 ? "x2  " : "bf16", sh.name,
                     pair ? " pair" : "     ", bn, sk, work, us, err_h, mag, err_st, bad_hi, cnt_bad);
              fflush(stdout);
            }
      }
    return 0;
  }

  // ---- layernorm ----
  for (int rows : {32, 256, B, 4 * B}) {
    float* hh = (float*)dmalloc((size_t)rows * d * 4);
    bf16* aa = (bf16*)dmalloc((size_t)rows * d * 2);
    ActOut y; y.hi = aa;
    printf("layernorm rows=%5d d=%d: %.2f us\n", rows, d, time_loop(st, 200, [&](int) { OK(launch_layernorm(hh, d, lnw, lnb, y, rows, d, st)); }));
    cudaFree(hh); cudaFree(aa);
  }
  // ---- GEMMs (decode shapes), weights rotate over 12 sets ----
  struct Shape { const char* name; int N, K; std::vector<bf16*>* w; int epi; bool res; };
  Shape shapes[] = {{"qkv", 3 * d, d, &w_qkv, EPI_NONE, false}, {"proj", d, d, &w_proj, EPI_RESIDUAL, true},
                    {"fc", 4 * d, d, &w_fc, EPI_GELU, false}, {"fc2", d, 4 * d, &w_fc2, EPI_RESIDUAL, true}};
  float2* stats = (float2*)dmalloc((size_t)64 * B * 8);  // [parts][B]
  float* colsum = (float*)dmalloc((size_t)V * 4);
  float* bias_v = (float*)dmalloc((size_t)V * 4);
  float* skws = (float*)dmalloc((size_t)4 * B * d * 4);
  int* skcnt = (int*)dmalloc(4096 * 4);
  // fused: 0 plain, 1 LayerNorm folded / statistics out (the engine's kernels), 2 the same as CTA pairs (cta_group::2)
  for (int fused = 0; fused < 3; ++fused)
    for (auto& s : shapes) {
      int bn_pick = 0, pair_pick = 0;
      gemm_bf16_pick(B, s.N, s.K, 0, 1, &bn_pick, fused == 2 ? &pair_pick : nullptr);
      for (int bn : {64, 128, 192, 256}) {
        if (fused == 1 && bn != bn_pick) continue;
        const int pair = fused == 2;
        std::vector<GemmBf16Args> args(SETS);
        for (int i = 0; i < SETS; ++i) {
          GemmBf16Args& g = args[i];
          OK(make_tma_2d_bf16(&g.a_hi, a, B, s.K, s.K, 128));
          OK(make_tma_2d_bf16(&g.w_hi, (*s.w)[i], s.N, s.K, s.K, pair ? bn / 2 : bn));
          g.M = B; g.N = s.N; g.K = s.K; g.block_n = bn; g.epilogue = s.epi; g.bias = bias; g.ld_out = s.N; g.pair = pair; g.w_static = 1;
          if (s.res) g.out.f32 = h; else g.out.hi = (s.N == 3 * d ? qkv : o);
          if (fused) {  // LayerNorm folded: statistics in for qkv / fc, bf16 copy + statistics out for the residual GEMMs
            if (s.res) { g.out.hi = o; g.stats_out = stats; g.ln_stats_ld = B; }
            else { g.ln_stats = stats; g.ln_parts = d / 32; g.ln_stats_ld = B; g.ln_colsum = colsum; }
          }
        }
        float us = time_loop(st, 240, [&](int i) { OK(launch_gemm_bf16(args[i % SETS], st)); });
        printf("gemm %-4s%s M=%d N=%d K=%d block_n=%3d: %7.2f us  %6.1f TFLOP/s  (picked %d%s)\n", s.name, fused == 2 ? " +LN pair" : fused ? " +LN     " : "         ", B,
               s.N, s.K, bn, us, 2.0 * B * s.N * s.K / us * 1e-6, bn_pick, pair_pick ? " pair" : "");
      }
    }
  // prefill-size GEMMs (M = 10 B rows), single CTAs vs pairs
  {
    const int Mp = 10 * B;
    bf16* ap = (bf16*)dmalloc((size_t)Mp * 4 * d * 2);
    bf16* op = (bf16*)dmalloc((size_t)Mp * 4 * d * 2);
    float* hp = (float*)dmalloc((size_t)Mp * d * 4);
    float2* statsp = (float2*)dmalloc((size_t)64 * Mp * 8);
    for (auto& s : shapes)
      for (int pair = 0; pair < 2; ++pair)
        for (int bn : {128, 192, 256}) {
          GemmBf16Args g;
          OK(make_tma_2d_bf16(&g.a_hi, ap, Mp, s.K, s.K, 128));
          OK(make_tma_2d_bf16(&g.w_hi, (*s.w)[0], s.N, s.K, s.K, pair ? bn / 2 : bn));
          g.M = Mp; g.N = s.N; g.K = s.K; g.block_n = bn; g.epilogue = s.epi; g.bias = bias; g.ld_out = s.N; g.pair = pair;
          if (s.res) { g.out.f32 = hp; g.out.hi = op; g.stats_out = statsp; g.ln_stats_ld = Mp; }
          else { g.out.hi = op; g.ln_stats = statsp; g.ln_parts = d / 32; g.ln_stats_ld = Mp; g.ln_colsum = colsum; }
          float us = time_loop(st, 40, [&](int) { OK(launch_gemm_bf16(g, st)); });
          printf("prefill gemm %-4s%s M=%d N=%d K=%d block_n=%3d: %7.2f us  %6.1f TFLOP/s\n", s.name, pair ? " pair" : "     ", Mp, s.N, s.K, bn, us,
                 2.0 * Mp * s.N * s.K / us * 1e-6);
        }
  }
  for (int pair = 0; pair < 2; ++pair)
    for (int bn : {128, 192, 256}) {
      GemmBf16Args g;
      OK(make_tma_2d_bf16(&g.a_hi, a, B, d, d, 128));
      OK(make_tma_2d_bf16(&g.w_hi, wte, V, d, d, pair ? bn / 2 : bn));
      g.M = B; g.N = V; g.K = d; g.block_n = bn; g.part_val = pv; g.part_idx = pi; g.part_ld = 2048; g.pair = pair;
      float us = time_loop(st, 50, [&](int) { OK(launch_gemm_bf16(g, st)); });
      printf("lm_head%s M=%d N=%d K=%d block_n=%d: %.2f us  %.1f TFLOP/s\n", pair ? " pair" : "     ", B, V, d, bn, us, 2.0 * B * V * d / us * 1e-6);
    }
  // ---- decode attention: ring geometries of the bulk-copy kernel (variant 0 = product) ----
  for (int variant : {0, 10, 12}) {
    attn_decode_set_variant(variant);
    for (int ctx : {11, 25, 39}) {
      int pos = ctx - 1;
      CK(cudaMemcpy(dpos, &pos, 4, cudaMemcpyHostToDevice));
      ActOut y; y.hi = o;
      float us = time_loop(st, 240, [&](int i) { OK(launch_attn_decode<bf16>(qkv, kc[i % SETS], vc[i % SETS], y, dpos, B, H, t_max, st)); });
      const double bytes = 2.0 * B * d * (2.0 * ctx + 2 + 3 + 1);
      printf("attn_decode v%d rows=%d ctx=%d: %.2f us  %.0f GB/s\n", variant, B, ctx, us, bytes / us * 1e-3);
    }
  }
  attn_decode_set_variant(0);
  // ---- what a kernel of this size can reach at all: plain streaming read of the same number of bytes (one launch each) ----
  const size_t stream_set = (size_t)160 << 20;
  uint8_t* stream_buf = (uint8_t*)dmalloc(stream_set * SETS);
  for (int ctx : {11, 25, 39}) {
    const size_t bytes = (size_t)(2.0 * B * d * (2.0 * ctx + 2 + 3 + 1));
    const size_t n16 = bytes / 16;
    float us = time_loop(st, 240, [&](int i) {
      stream_read_kernel<<<148 * 4, 512, 0, st>>>((const uint4*)(stream_buf + (size_t)(i % SETS) * stream_set), n16, (unsigned*)pi);
    });
    printf("stream_read %zu MB (ctx=%d equivalent): %.2f us  %.0f GB/s\n", bytes >> 20, ctx, us, bytes / us * 1e-3);
  }
  return 0;
}
