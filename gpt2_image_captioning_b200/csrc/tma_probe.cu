// Hardware probe (not part of the library): how fast can one SM / the whole chip pull operand tiles through a TMA ->
// mbarrier ring, as a function of bytes in flight, box size and source (L2-resident vs HBM)?  No MMA: the consumer
// thread releases each stage as soon as it lands, so the numbers are the ceiling of any TMA-fed GEMM main loop.
//   make tma_probe && build/tma_probe
#include <stdarg.h>
#include <stdlib.h>

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
bool pdl_enabled() { return false; }
}  // namespace gic

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct alignas(64) ProbeParams {
  TmaDesc map;        // [K/64 atoms][rows][64] bf16, box = box_rows x 64 = box_rows * 128 bytes
  int box_rows;       // rows per TMA
  int tmas_per_stage; // 1..4
  int stages;
  int iters;          // stages streamed per CTA
  int rows_total;     // source rows (row = 768 bf16 = 6 k-blocks of 128)
  int shared_source;  // 1: every CTA walks the same rows (L2 hits after the first toucher); 0: disjoint rows per CTA
  long long* cycles;  // [grid] elapsed SM clocks of the consumer
};

__global__ void __launch_bounds__(64, 1) tma_stream_probe(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.box_rows * 128 * p.tmas_per_stage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int row_blocks = p.rows_total / p.box_rows;  // boxes along the row axis
  if (warp == 0 && lane == 0) {
    // box index walks (row block, k-block) like a GEMM operand stream: 12 k-blocks of 64 per row block.  All indices advance
    // incrementally: an integer division in this loop would cost more than the TMA issue it is trying to measure.
    int kb = 0, rb = p.shared_source ? (int)((blockIdx.x * 7) % row_blocks) : (int)(((long)blockIdx.x * p.iters * p.tmas_per_stage / 12) % row_blocks);
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&empty_bar[s], ph ^ 1);
      mbar_expect_tx(&full_bar[s], stage_bytes);
      for (int t = 0; t < p.tmas_per_stage; ++t) {
        tma_load_2d(smem + (size_t)s * stage_bytes + (size_t)t * p.box_rows * 128, &p.map, &full_bar[s], kb * 64, rb * p.box_rows);
        if (++kb == 12) { kb = 0; if (++rb == row_blocks) rb = 0; }
      }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    const long long c0 = clock64();
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      mbar_arrive(&empty_bar[s]);
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    p.cycles[blockIdx.x] = clock64() - c0;
  }
}

int main() {
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  OK(tma_init());
  const int K = 768;
  const size_t rows_big = (size_t)1 << 20;  // 1M rows x 768 x 2 B = 1.5 GB (HBM stream)
  bf16* buf; CK(cudaMalloc(&buf, rows_big * K * 2)); CK(cudaMemset(buf, 0, rows_big * K * 2));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(tma_stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("%-7s %5s %8s %5s %6s %9s | %9s %10s %10s %12s\n", "source", "grid", "box_rows", "tmas", "stages", "inflightKB", "us", "GB/s/SM", "TB/s chip",
         "cyc/stage");
  struct Cfg { int box_rows, tmas, stages; };
  const Cfg cfgs[] = {{128, 1, 6}, {128, 2, 3}, {128, 1, 3}, {128, 2, 2}, {256, 1, 3}, {64, 1, 12}, {64, 2, 6}, {64, 1, 6}, {64, 1, 3},
                      {128, 1, 2}, {256, 1, 2}, {192, 1, 4}};
  for (int shared_source = 1; shared_source >= 0; --shared_source)
    for (int grid : {1, 16, 74, 148})
      for (const Cfg& c : cfgs) {
        ProbeParams p;
        const size_t rows = shared_source ? 8192 : rows_big;  // 8192 rows = 12.6 MB: L2 resident
        OK(make_tma_2d_bf16(&p.map, buf, rows, K, K, c.box_rows));
        p.box_rows = c.box_rows; p.tmas_per_stage = c.tmas; p.stages = c.stages; p.rows_total = (int)rows; p.shared_source = shared_source;
        const int stage_bytes = c.box_rows * 128 * c.tmas;
        const size_t target = (size_t)(grid == 1 ? 64 : 24) << 20;  // bytes per CTA
        p.iters = (int)(target / stage_bytes);
        if (!shared_source && (size_t)p.iters * grid * c.tmas * c.box_rows > rows_big * 12) p.iters = (int)(rows_big * 12 / ((size_t)grid * c.tmas * c.box_rows));
        p.cycles = cyc;
        const size_t smem = (size_t)c.stages * stage_bytes + 2048;
        for (int rep = 0; rep < 2; ++rep) {  // first pass warms L2 / instruction cache
          CK(cudaEventRecord(e0, st));
          tma_stream_probe<<<grid, 64, smem, st>>>(p);
          CK(cudaEventRecord(e1, st));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes_cta = (double)p.iters * stage_bytes;
        long long hc[148]; CK(cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost));
        double cmax = 0; for (int i = 0; i < grid; ++i) cmax = hc[i] > cmax ? hc[i] : cmax;
        printf("%-7s %5d %8d %5d %6d %9d | %9.1f %10.1f %10.2f %12.0f\n", shared_source ? "L2" : "HBM", grid, c.box_rows, c.tmas, c.stages,
               c.stages * stage_bytes / 1024, ms * 1e3, bytes_cta / (ms * 1e-3) / 1e9, bytes_cta * grid / (ms * 1e-3) / 1e12, cmax / p.iters);
      }
  return 0;
}
