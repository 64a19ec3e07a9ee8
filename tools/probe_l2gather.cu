// Ceiling probe for the table gather of the VE kernels: how many RANDOM 32-byte sectors per second can a B200 fetch from an
// L2-resident table?  Every thread reads ILP independent sectors per iteration (index stream read coalesced, 4 B per
// access) and writes 4 B per access, like a MAP row.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_l2gather tools/probe_l2gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int ILP, int BYTES>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ idx, int64_t n, const float4* __restrict__ table,
                                                float* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i + (ILP - 1) * stride < n; i += ILP * stride) {
    uint32_t j[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) j[k] = __ldcs(idx + i + k * stride);
    float4 v[ILP], w[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      v[k] = __ldg(table + 2 * size_t(j[k]));
      if (BYTES == 32) w[k] = __ldg(table + 2 * size_t(j[k]) + 1);
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      float s = v[k].x + v[k].y + v[k].z + v[k].w;
      if (BYTES == 32) s += w[k].x + w[k].y + w[k].z + w[k].w;
      __stcs(out + i + k * stride, s);
    }
  }
}

// the shapes of the VE gather kernels: EW streamed 32-bit words in per access (the evidence codes), ONE gather of GB bytes
// (8: MAP on a binary target, 16: a 4-valued target, 32: four fused binary targets) and SB bytes stored per access
template <int ILP, int EW, int GB, int SB>
__global__ void __launch_bounds__(256) k_shape(const uint32_t* __restrict__ idx, int64_t n, const float* __restrict__ table,
                                               float* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i + (ILP - 1) * stride < n; i += ILP * stride) {
    uint32_t j[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      j[k] = __ldcs(idx + i + k * stride);
#pragma unroll
      for (int e = 1; e < EW; ++e) j[k] ^= __ldcs(idx + e * n + i + k * stride) & 0u;      // extra evidence words (all read)
    }
    float v[ILP][8];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      const float* p = table + 8 * size_t(j[k]);
      if (GB == 32) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[k][0]), "=f"(v[k][1]), "=f"(v[k][2]), "=f"(v[k][3]), "=f"(v[k][4]), "=f"(v[k][5]), "=f"(v[k][6]), "=f"(v[k][7])
                     : "l"(p));
      } else if (GB == 16) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
      } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[k][0] = t.x; v[k][1] = t.y;
      }
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      float* o = out + (i + k * stride) * (SB / 4);
      if (SB == 32) {
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "f"(v[k][0]), "f"(v[k][1]), "f"(v[k][2]), "f"(v[k][3]),
                     "f"(v[k][4]), "f"(v[k][5]), "f"(v[k][6]), "f"(v[k][7]) : "memory");
      } else if (SB == 16) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(v[k][0], v[k][1], v[k][2], v[k][3]));
      } else {
        __stcs(o, v[k][0] + v[k][1]);
      }
    }
  }
}

template <int ILP, int EW, int GB, int SB>
static void run_shape(const char* name, const uint32_t* idx, int64_t n, const float* table, float* out, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 8;
  for (int w = 0; w < 3; ++w) k_shape<ILP, EW, GB, SB><<<blocks, 256>>>(idx, n, table, out);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) k_shape<ILP, EW, GB, SB><<<blocks, 256>>>(idx, n, table, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double s = ms / 1e3 / reps;
  printf("%-34s n %9lld  ilp %d  in %2d B  gather %2d B  out %2d B  %8.1f us  %7.2f G rows/s  %7.1f GB/s algorithmic\n", name, (long long)n, ILP,
         4 * EW, GB, SB, s * 1e6, n / s / 1e9, n * double(4 * EW + SB) / s / 1e9);
}

template <int ILP, int BYTES>
static void run(const char* name, const uint32_t* idx, int64_t n, const float4* table, float* out, int ctas_per_sm, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * ctas_per_sm;
  for (int w = 0; w < 3; ++w) k_gather<ILP, BYTES><<<blocks, 256>>>(idx, n, table, out);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_gather<ILP, BYTES><<<blocks, 256>>>(idx, n, table, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double s = ms / 1e3 / reps;
  printf("%-28s ilp %d  %d B/access  ctas/sm %d  %8.1f us  %7.2f G sectors/s  (stream %6.1f GB/s)\n", name, ILP, BYTES, ctas_per_sm,
         s * 1e6, n / s / 1e9, n * 8.0 / s / 1e9);
}

int main(int argc, char** argv) {
  const int64_t n = int64_t(1) << 24;                       // accesses per launch (the Alarm batch has 2^24 rows)
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (double mb : {0.84, 6.7, 16.0, 27.0, 64.0, 512.0}) {
    const size_t sectors = size_t(mb * 1e6 / 32);
    float4* table; uint32_t* idx; float* out;
    cudaMalloc(&table, sectors * 32); cudaMemset(table, 0, sectors * 32);
    cudaMalloc(&idx, n * 4); cudaMalloc(&out, n * 4);
    std::vector<uint32_t> h(n);
    uint64_t z = 88172645463325252ull;
    for (int64_t i = 0; i < n; ++i) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; h[i] = uint32_t(z % sectors); }
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    char name[64];
    snprintf(name, sizeof name, "table %.1f MB", mb);
    run<1, 16>(name, idx, n, table, out, 8, sms);
    run<4, 16>(name, idx, n, table, out, 8, sms);
    run<8, 16>(name, idx, n, table, out, 8, sms);
    run<4, 32>(name, idx, n, table, out, 8, sms);
    run<8, 16>(name, idx, n, table, out, 4, sms);
    cudaFree(table); cudaFree(idx); cudaFree(out);
  }
  {
    // kernel shapes (table L2-resident, 27 MB like the fused Alarm table / 16 MB like a 200-node pattern / 6.7 MB MAP)
    const int64_t nmax = int64_t(1) << 24;
    const size_t sectors = size_t(27e6 / 32);
    float* table; uint32_t* idx; float* out;
    cudaMalloc(&table, sectors * 32); cudaMemset(table, 0, sectors * 32);
    cudaMalloc(&idx, nmax * 4 * 3); cudaMalloc(&out, nmax * 32);
    std::vector<uint32_t> h(nmax * 3);
    uint64_t z = 88172645463325252ull;
    for (int64_t i = 0; i < nmax * 3; ++i) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; h[i] = uint32_t(z % (sectors / 2)); }
    cudaMemcpy(idx, h.data(), nmax * 12, cudaMemcpyHostToDevice);
    run_shape<4, 3, 32, 32>("alarm x4 targets (headline)", idx, nmax, table, out, sms);
    run_shape<2, 3, 32, 32>("alarm x4 targets (headline)", idx, nmax, table, out, sms);
    run_shape<4, 3, 16, 16>("200-node pattern", idx, nmax, table, out, sms);
    run_shape<4, 3, 16, 16>("200-node pattern, 1M-row launch", idx, int64_t(1) << 20, table, out, sms);
    run_shape<2, 3, 16, 16>("200-node pattern, 1M-row launch", idx, int64_t(1) << 20, table, out, sms);
    run_shape<4, 3, 8, 4>("alarm MAP", idx, nmax, table, out, sms);
    run_shape<8, 3, 8, 4>("alarm MAP", idx, nmax, table, out, sms);
  }
  return 0;
}
