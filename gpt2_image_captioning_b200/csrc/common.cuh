// Shared helpers for the sm_100a caption-generation kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gic_b200.h"

namespace gic {

// ---- error plumbing: never throw across the C ABI; record a message, return a code --------------
void set_error(const char* fmt, ...);
const char* get_error();
void note_launch();
struct StepTrace;
StepTrace trace_desc();             // descriptor for the NEXT launch (a fresh slot per call; null buffer: off), passed to the kernel by value
bool trace_on();
unsigned int trace_generation();   // bumped by every install: a captured graph holds the descriptor it was captured with  // every kernel launch of the library reports here (gic_launch_count)

#define GIC_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::gic::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return GIC_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define GIC_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::gic::set_error(__VA_ARGS__);           \
      return GIC_ERR_INVALID;                  \
    }                                          \
  } while (0)

#define GIC_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != GIC_OK) return _r; \
  } while (0)

// ---- launches: programmatic dependent launch (PDL) ------------------------------------------------
// A decode step is ~90 short dependent kernels; with plain stream order each pays the full launch gap (3.65 us measured
// per back-to-back empty launch on this box).  With PDL the next kernel's CTAs are scheduled as soon as every CTA of the
// current one has started (griddepcontrol.launch_dependents at the top of each kernel) and run their prologue (barrier
// init, TMEM alloc, descriptor prefetch) while it finishes; they touch global memory only after griddepcontrol.wait,
// which returns when all earlier grids have completed and flushed.  GIC_NO_PDL=1 turns the launch attribute off.
bool pdl_enabled();
// Persistent kernels (GEMM, decode attention) size their grids to min(work, cta_limit()): the whole device unless the engine is
// issuing one of several concurrent row-group chains, each of which is given its share of the SMs (set_cta_limit(0) = all).
int cta_limit();
void set_cta_limit(int ctas);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

typedef __nv_bfloat16 bf16;

// ---- in-situ step timeline (tools/step_timeline.py) ---------------------------------------------------------------
// Thread 0 of EVERY block of the decode-step kernels stamps %globaltimer after griddepcontrol.wait and at its end; the launch's
// record (kind, begin_ns = min over blocks, end_ns = max over blocks) lives in slot `slot` of a device buffer, the slot being
// handed out by the host per launch (trace_desc(): null buffer = tracing off, the launch carries a null pointer and pays
// nothing).  The tool pre-fills begin with ~0 and end with 0.  A captured graph keeps the slots it was captured with, so the
// traced run captures all decode steps into one graph.
struct StepTrace { unsigned long long* buf; unsigned int slot; unsigned int cap; };
enum TraceKind { TRACE_GEMM = 1, TRACE_ATTN_DECODE = 2, TRACE_LAYERNORM = 3, TRACE_FINALIZE = 4, TRACE_ATTN_PREFILL = 5, TRACE_RESCORE = 6 };
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) : : "memory");  // (memory clobber: must not drift across barriers / waits)
  return t;
}
__device__ __forceinline__ int trace_begin(const StepTrace& tr, int kind, int detail) {
  if (threadIdx.x != 0 || tr.buf == nullptr || tr.slot >= tr.cap) return -1;
  if (blockIdx.x == 0) tr.buf[3 * tr.slot] = ((unsigned long long)(unsigned int)detail << 8) | (unsigned long long)kind;
  atomicMin(tr.buf + 3 * tr.slot + 1, globaltimer_ns());
  return (int)tr.slot;
}
__device__ __forceinline__ void trace_end(const StepTrace& tr, int slot) {
  if (slot >= 0) atomicMax(tr.buf + 3 * slot + 2, globaltimer_ns());
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// GELU, tanh form -- NewGELUActivation (HF:activations.py:66)
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x * x * x)));
}

// the same function as x * sigmoid(2 u), u = sqrt(2/pi) (x + 0.044715 x^3): 0.5 x (1 + tanh u) = x / (1 + e^(-2u)).  One ex2 + one
// reciprocal instead of tanhf's ~25 instructions, relative error ~3e-7 (fp32 rounding of 1 + e; ex2.approx is good to 2^-22) -- the
// bf16x2 engine's GELU, whose result is then split into bf16 hi + lo (2^-17)
__device__ __forceinline__ float gelu_tanh_sigmoid(float x) {
  const float k2 = 2.0f * 0.7978845608028654f;
  return __fdividef(x, 1.0f + __expf(-k2 * (x + 0.044715f * x * x * x)));
}

// same with the single-instruction MUFU.TANH (abs err ~5e-4): used where the result is stored as bf16 anyway
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k = 0.7978845608028654f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k * (x + 0.044715f * x * x * x)));
  return 0.5f * x * (1.0f + t);
}

// 16-byte vector of activations, unpacked to floats
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 v = __bfloat1622float2(h[i]);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gic
