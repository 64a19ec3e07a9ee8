// Micro-benchmark: lane-private (bank-conflict-free) shared-memory counters against shared atomics, for the count kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_lanepriv tools/probe_lanepriv.cu
// Every warp owns one 256-cell table.  Layouts:
//   mode 0  shared uint32 table [256], atomicAdd(+1) at a random cell          (ATOMS.POPC.INC, bank conflicts)
//   mode 1  lane-private byte counters packed 4 per word: word (cell>>2)*32+lane, red.shared.add of 1<<(8*(cell&3))
//   mode 2  same layout, plain byte read-modify-write (ld.shared.u8 / st.shared.u8), no atomics
//   mode 3  same layout, 32-bit read-modify-write of the word
//   mode 4  shared uint32 table, red.shared.add with a register value (non-POPC form), random cell
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int TPB>
__global__ void __launch_bounds__(TPB) k(int iters, unsigned long long* sink) {
  extern __shared__ __align__(16) uint32_t tbl[];
  constexpr int WORDS_PER_WARP = (MODE == 0 || MODE == 4) ? 256 : 64 * 32;
  for (int i = threadIdx.x; i < WORDS_PER_WARP * (TPB / 32); i += TPB) tbl[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* my = tbl + warp * WORDS_PER_WARP;
  const uint32_t base = smem_u32(my);
  uint32_t s = hash32(blockIdx.x * TPB + threadIdx.x + 1);
  unsigned long long total = 0;
  for (int it = 0; it < iters; ++it) {
    // 8 updates per iteration (the count kernel handles 8 samples per lane per step)
    uint32_t c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { s = s * 1664525u + 1013904223u; c[u] = (s >> 10) & 255u; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t cell = c[u];
      if (MODE == 0) {
        atomicAdd(&my[cell], 1u);
      } else if (MODE == 4) {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + cell * 4), "r"(1u + (cell >> 31)) : "memory");
      } else if (MODE == 1) {
        const uint32_t addr = base + ((cell >> 2) * 32 + lane) * 4;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << (8 * (cell & 3))) : "memory");
      } else if (MODE == 2) {
        const uint32_t addr = base + ((cell >> 2) * 32 + lane) * 4 + (cell & 3);
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v + 1) : "memory");
      } else {
        const uint32_t addr = base + ((cell >> 2) * 32 + lane) * 4;
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v + (1u << (8 * (cell & 3)))) : "memory");
      }
    }
    if (MODE != 0 && MODE != 4 && (it % 31) == 30) {
      // flush: byte counters would overflow after 255 updates per lane (31 iterations x 8 = 248): sum and clear
      __syncwarp();
      for (int w = lane; w < WORDS_PER_WARP / 4; w += 32) {
        uint4 v = reinterpret_cast<uint4*>(my)[w];
        total += __dp4a(v.x, 0x01010101u, 0u) + __dp4a(v.y, 0x01010101u, 0u) + __dp4a(v.z, 0x01010101u, 0u) + __dp4a(v.w, 0x01010101u, 0u);
        reinterpret_cast<uint4*>(my)[w] = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WORDS_PER_WARP * (TPB / 32); i += TPB) {
    const uint32_t v = tbl[i];
    total += (MODE == 0 || MODE == 4) ? v : __dp4a(v, 0x01010101u, 0u);
  }
  atomicAdd(sink, total);
}

template <int MODE, int TPB>
void run(const char* name, int ctas_per_sm) {
  int sms = 148, iters = 31 * 40;
  unsigned long long* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  size_t smem = size_t((MODE == 0 || MODE == 4) ? 256 : 64 * 32) * 4 * (TPB / 32);
  cudaFuncSetAttribute(k<MODE, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  k<MODE, TPB><<<sms * ctas_per_sm, TPB, smem>>>(31, sink);
  cudaMemset(sink, 0, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE, TPB><<<sms * ctas_per_sm, TPB, smem>>>(iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  unsigned long long got = 0; cudaMemcpy(&got, sink, 8, cudaMemcpyDeviceToHost);
  double updates = double(sms) * ctas_per_sm * TPB * double(iters) * 8;
  cudaError_t e = cudaGetLastError();
  printf("%-44s tpb=%d ctas/sm=%d smem=%6zu %8.1f G upd/s (%.2f upd/clk/SM @1.9GHz) sum %s %s\n", name, TPB, ctas_per_sm, smem,
         updates / ms / 1e6, updates / ms / 1e6 / 148 / 1.9, got == (unsigned long long)updates ? "ok" : "WRONG",
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(sink);
}

int main() {
  run<0, 256>("shared table, atomicAdd +1 (POPC.INC)", 1);
  run<0, 256>("shared table, atomicAdd +1 (POPC.INC)", 2);
  run<0, 512>("shared table, atomicAdd +1 (POPC.INC)", 1);
  run<4, 512>("shared table, red.add reg value", 1);
  run<1, 256>("lane-private bytes, red.add 1<<8k", 1);
  run<1, 256>("lane-private bytes, red.add 1<<8k", 2);
  run<1, 512>("lane-private bytes, red.add 1<<8k", 1);
  run<2, 256>("lane-private bytes, ld.u8/st.u8", 1);
  run<2, 256>("lane-private bytes, ld.u8/st.u8", 2);
  run<2, 512>("lane-private bytes, ld.u8/st.u8", 1);
  run<3, 256>("lane-private bytes, ld.u32/st.u32", 1);
  run<3, 256>("lane-private bytes, ld.u32/st.u32", 2);
  run<3, 512>("lane-private bytes, ld.u32/st.u32", 1);
  return 0;
}
