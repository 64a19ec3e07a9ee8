// Throughput of red.global.add.u32 on L2-resident tables (random cells), as a possible second lane for CPT counting beside
// the shared-memory atomics.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_l2atoms tools/probe_l2atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k(unsigned* __restrict__ tbl, int cells_mask, int n_tables, long long per_thread, int same_warp_cell) {
  uint64_t x = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  for (long long i = 0; i < per_thread; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    unsigned cell = unsigned(x) & cells_mask;
    unsigned t = unsigned(x >> 40) % n_tables;
    if (same_warp_cell) cell = __shfl_sync(0xffffffffu, cell, 0);
    asm volatile("red.global.add.u32 [%0], 1;" ::"l"(tbl + size_t(t) * (cells_mask + 1) + cell) : "memory");
  }
}

int main() {
  int sm = 148;
  unsigned* tbl;
  cudaMalloc(&tbl, 64 << 20);
  cudaMemset(tbl, 0, 64 << 20);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const long long per_thread = 2000;
  for (int same = 0; same < 2; ++same)
    for (int cells : {256, 1024, 65536})
      for (int tables : {1, 16, 200}) {
        int blocks = sm * 8;
        k<<<blocks, 256>>>(tbl, cells - 1, tables, 100, same);
        cudaEventRecord(a);
        k<<<blocks, 256>>>(tbl, cells - 1, tables, per_thread, same);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double ups = double(blocks) * 256 * per_thread / (ms * 1e-3);
        printf("%s cells=%6d tables=%3d  %8.1f G updates/s (%.2f per clk per SM @1.9GHz)\n", same ? "warp-uniform cell" : "random cell      ", cells, tables, ups / 1e9, ups / 148 / 1.9e9);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
