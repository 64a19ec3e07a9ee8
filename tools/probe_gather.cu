// Floor probe for the Asia-sized gather (1M rows, 4 evidence columns, 3 targets x float2): how fast can ANY kernel
// of this shape go on a B200?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_gather tools/probe_gather.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROWS = 1 << 20;
constexpr int RING = 9;

template <int RPT_WORDS, bool SMEM_LUT, int STORE = 0>
__global__ void __launch_bounds__(256) k_min(const uint8_t* __restrict__ ev, int64_t ld, int64_t n_words, const float2* __restrict__ lut,
                                            float2* __restrict__ o0, float2* __restrict__ o1, float2* __restrict__ o2) {
  __shared__ float2 s[96];
  if (SMEM_LUT) {
    if (threadIdx.x < 96) s[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
  }
  const float2* L = SMEM_LUT ? s : lut;
  for (int64_t q = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * RPT_WORDS; q < n_words; q += int64_t(gridDim.x) * blockDim.x * RPT_WORDS) {
    uint32_t w[4][RPT_WORDS];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (RPT_WORDS == 4) {
        uint4 v = *reinterpret_cast<const uint4*>(ev + c * ld + q * 4);
        w[c][0] = v.x; w[c][1] = v.y; w[c][2] = v.z; w[c][3] = v.w;
      } else {
#pragma unroll
        for (int k = 0; k < RPT_WORDS; ++k) w[c][k] = *reinterpret_cast<const uint32_t*>(ev + c * ld + (q + k) * 4);
      }
    }
#pragma unroll
    for (int k = 0; k < RPT_WORDS; ++k) {
      const uint32_t acc = w[0][k] * 8u + w[1][k] * 4u + w[2][k] * 2u + w[3][k];
      float2 r[3][4];
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) r[t][j] = L[t * 32 + ((acc >> (8 * j)) & 0xff)];
      float2* outs[3] = {o0, o1, o2};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        float4* dst = reinterpret_cast<float4*>(outs[t] + (q + k) * 4);
        if (STORE == 0) {
          dst[0] = make_float4(r[t][0].x, r[t][0].y, r[t][1].x, r[t][1].y);
          dst[1] = make_float4(r[t][2].x, r[t][2].y, r[t][3].x, r[t][3].y);
        } else if (STORE == 1) {   // one 256-bit store per thread
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(r[t][0].x), "f"(r[t][0].y), "f"(r[t][1].x),
                       "f"(r[t][1].y), "f"(r[t][2].x), "f"(r[t][2].y), "f"(r[t][3].x), "f"(r[t][3].y) : "memory");
        } else {                   // no write at all (read + compute only): lower bound of the non-store part
          if (r[t][0].x == 123.456f) dst[0] = make_float4(0, 0, 0, 0);
        }
      }
    }
  }
}

template <typename F>
float time_launches(F launch, int iters) {
  for (int i = 0; i < 20; ++i) launch(i);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) launch(i);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters * 1e3f;
}

int main() {
  int64_t ld = ROWS;
  std::vector<uint8_t*> ev(RING); std::vector<float2*> o(RING * 3);
  for (int r = 0; r < RING; ++r) {
    cudaMalloc(&ev[r], 4 * ld); cudaMemset(ev[r], 1, 4 * ld);
    for (int t = 0; t < 3; ++t) cudaMalloc(&o[r * 3 + t], ROWS * sizeof(float2));
  }
  float2* lut; cudaMalloc(&lut, 96 * sizeof(float2)); cudaMemset(lut, 0, 96 * sizeof(float2));
  const int64_t n_words = ROWS / 4;
  const double bytes = ROWS * 28.0;
  auto report = [&](const char* name, float us) { printf("%-46s %7.2f us/launch  %7.1f GB/s  %.3f of 6521\n", name, us, bytes / us / 1e3, bytes / us / 1e3 / 6521.1); };
  for (int grid : {1024, 888, 592, 296}) {
    char nm[128];
    snprintf(nm, sizeof nm, "1 word/thread, smem LUT, grid %d", grid);
    report(nm, time_launches([&](int i) { int r = i % RING; k_min<1, true><<<grid, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
    snprintf(nm, sizeof nm, "1 word/thread, L1 LUT, grid %d", grid);
    report(nm, time_launches([&](int i) { int r = i % RING; k_min<1, false><<<grid, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  }
  report("1 word/thread, smem LUT, 256-bit stores, grid 1024", time_launches([&](int i) { int r = i % RING; k_min<1, true, 1><<<1024, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  report("1 word/thread, L1 LUT, 256-bit stores, grid 1024", time_launches([&](int i) { int r = i % RING; k_min<1, false, 1><<<1024, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  report("1 word/thread, L1 LUT, 256-bit stores, grid 592", time_launches([&](int i) { int r = i % RING; k_min<1, false, 1><<<592, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  report("1 word/thread, smem LUT, NO stores, grid 1024", time_launches([&](int i) { int r = i % RING; k_min<1, true, 2><<<1024, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  report("empty-ish kernel (n_words=0), grid 1024", time_launches([&](int i) { int r = i % RING; k_min<1, true, 2><<<1024, 256>>>(ev[r], ld, 0, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  for (int grid : {256, 148, 296}) {
    char nm[128];
    snprintf(nm, sizeof nm, "4 words/thread (128-bit loads), smem LUT, grid %d", grid);
    report(nm, time_launches([&](int i) { int r = i % RING; k_min<4, true><<<grid, 256>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]); }, 300));
  }
  // same through a CUDA graph of one ring cycle
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int r = 0; r < RING; ++r) k_min<1, true><<<1024, 256, 0, s>>>(ev[r], ld, n_words, lut, o[r*3], o[r*3+1], o[r*3+2]);
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < 30; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  report("graph of ring cycle, 1 word/thread grid 1024", ms / (30 * RING) * 1e3f);
  // plain copy kernel-free reference: cudaMemcpyAsync D2D of 14 MB (read) + (write) = 28 MB traffic
  uint8_t *x, *y; cudaMalloc(&x, 14 << 20); cudaMalloc(&y, 14 << 20);
  report("cudaMemcpy D2D 14 MB (28 MB traffic)", time_launches([&](int) { cudaMemcpyAsync(y, x, 14 << 20, cudaMemcpyDeviceToDevice, 0); }, 300));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
