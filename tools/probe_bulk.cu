// Micro-benchmark: throughput of 1-D bulk async copies (cp.async.bulk global -> shared, mbarrier complete_tx) as a function
// of the copy size, with a fixed number of bytes per stage: does a tile made of many narrow column copies cost more than the
// same bytes in a few wide ones?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_bulk tools/probe_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// columns: n_cols arrays of `ld` bytes; a tile = `copy` bytes of each column; S stages; one warp issues (2 copies per lane max)
__global__ void __launch_bounds__(128) k(const uint8_t* __restrict__ base, int64_t ld, int n_cols, int copy, int64_t n_tiles, int stages, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[8];
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint32_t stage_bytes = uint32_t(n_cols) * copy;
  const int lane = threadIdx.x & 31;
  const int64_t my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t i) {
    const int s = int(i % stages);
    const int64_t tile = blockIdx.x + i * gridDim.x;
    if (lane == 0) mbar_expect_tx(&full[s], stage_bytes);
    __syncwarp();
    for (int c = lane; c < n_cols; c += 32) bulk_g2s(smem + size_t(s) * stage_bytes + size_t(c) * copy, base + int64_t(c) * ld + tile * copy, copy, &full[s]);
  };
  if (threadIdx.x < 32) for (int64_t i = 0; i < my && i < stages - 1; ++i) issue(i);
  unsigned acc = 0;
  for (int64_t i = 0; i < my; ++i) {
    const int s = int(i % stages);
    mbar_wait(&full[s], uint32_t(i / stages) & 1u);
    acc += smem[size_t(s) * stage_bytes + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 32 && i + stages - 1 < my) issue(i + stages - 1);
  }
  if (acc == 0xffffffffu) *sink = acc;
}

int main() {
  const int64_t total = int64_t(4) << 30;     // 4 GiB of source
  uint8_t* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  unsigned* sink; cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int cps : {1, 2}) for (int stage_kb : {32}) for (int stages : {3, 4}) for (int copy : {512, 1024, 2048, 4096, 8192, 16384}) {
    const int n_cols = stage_kb * 1024 / copy;
    if (n_cols > 64 || n_cols < 1) continue;
    const int64_t ld = total / n_cols / 16 * 16;
    const int64_t n_tiles = ld / copy;
    const size_t smem = size_t(stages) * stage_kb * 1024;
    if (smem * cps > 220 * 1024) continue;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148 * cps, 128, smem>>>(buf, ld, n_cols, copy, n_tiles / 8, stages, sink);
    cudaEventRecord(a);
    k<<<148 * cps, 128, smem>>>(buf, ld, n_cols, copy, n_tiles, stages, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double bytes = double(n_tiles) * n_cols * copy;
    printf("ctas/sm=%d stages=%d stage=%2d KB  copy=%5d B x %2d cols  %7.1f GB/s  (%.1f B/clk/SM, %.0f clk per copy) %s\n", cps, stages, stage_kb, copy, n_cols,
           bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9, copy / (bytes / ms / 1e6 / 148 / 1.9), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
