// Micro-benchmark: shared-memory atomic throughput on B200 for the histogram access patterns of the count kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_atoms tools/probe_atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

// mode 0: uniform random cell in [0,cells); mode 1: skewed (half of the mass on 1/16 of the cells);
// mode 2: conflict-free (lane-distinct banks); mode 3: all lanes same address
template <int MODE, bool MATCH>
__global__ void __launch_bounds__(256) atoms_kernel(int cells, int iters, unsigned long long* sink) {
  extern __shared__ uint32_t tbl[];
  for (int i = threadIdx.x; i < cells; i += blockDim.x) tbl[i] = 0;
  __syncthreads();
  uint32_t s = hash32(blockIdx.x * 256 + threadIdx.x + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int u = 0; u < 4; ++u) {
      s = s * 1664525u + 1013904223u;
      uint32_t r = s >> 8;
      uint32_t idx;
      if (MODE == 0) idx = r & (cells - 1);
      else if (MODE == 1) idx = (r & 1) ? (r >> 1) & (cells / 16 - 1) : (r >> 1) & (cells - 1);
      else if (MODE == 2) idx = ((r >> 5) & (cells / 32 - 1)) * 32 + (threadIdx.x & 31);
      else idx = (it * 4 + u) & (cells - 1);
      if (MATCH) {
        unsigned m = __match_any_sync(0xffffffffu, idx);
        if ((__ffs(m) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&tbl[idx], (uint32_t)__popc(m));
      } else {
        atomicAdd(&tbl[idx], 1u);
      }
    }
  }
  __syncthreads();
  unsigned long long t = 0;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) t += tbl[i];
  if (t == 0xffffffffffffull) *sink = t;
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, t);
}

template <int MODE, bool MATCH>
void run(const char* name, int cells, int ctas_per_sm) {
  int sms = 148, iters = 2000;
  unsigned long long* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  size_t smem = cells * 4;
  cudaFuncSetAttribute(atoms_kernel<MODE, MATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  atoms_kernel<MODE, MATCH><<<sms * ctas_per_sm, 256, smem>>>(cells, 10, sink);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  atoms_kernel<MODE, MATCH><<<sms * ctas_per_sm, 256, smem>>>(cells, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double updates = double(sms) * ctas_per_sm * 256.0 * iters * 4;
  cudaError_t e = cudaGetLastError();
  printf("%-34s cells=%6d ctas/sm=%d  %8.1f G updates/s  (%.2f updates/clk/SM @1.9GHz) %s\n", name, cells, ctas_per_sm,
         updates / ms / 1e6, updates / ms / 1e6 / 148 / 1.9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(sink);
}

int main() {
  for (int cps : {1, 2, 4, 8}) {
    run<0, false>("uniform random", 1024, cps);
    run<0, false>("uniform random", 16384, cps);
    run<1, false>("skewed", 1024, cps);
    run<2, false>("conflict-free banks", 1024, cps);
    run<3, false>("same address", 1024, cps);
    run<0, false>("uniform random tiny", 8, cps);
    run<0, false>("uniform random 64", 64, cps);
  }
  return 0;
}
