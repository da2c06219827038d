// Standalone micro-benchmark (development tool, not product): what does a 4-corner 256-byte-row gather +
// 256-byte store reach on this GPU as a function of kernel structure?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int W = 512, H = 256, C4 = 16, N = 40;
constexpr long long NPIX = (long long)N * H * W;

__device__ __forceinline__ float4 ldnc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 comb(float4 a, float4 b, float4 c, float4 d) {
  return make_float4(a.x * 0.4f + b.x * 0.3f + c.x * 0.2f + d.x * 0.1f, a.y * 0.4f + b.y * 0.3f + c.y * 0.2f + d.y * 0.1f,
                     a.z * 0.4f + b.z * 0.3f + c.z * 0.2f + d.z * 0.1f, a.w * 0.4f + b.w * 0.3f + c.w * 0.2f + d.w * 0.1f);
}
// clamp source pixel for the 4 corners of output pixel (n,i,j) shifted by (sx,sy)
__device__ __forceinline__ void corners(long long pix, int sx, int sy, long long o[4]) {
  const int j = pix % W;
  const long long r = pix / W;
  const int i = r % H;
  const long long n = r / H;
  const int x0 = min(max(j + sx, 0), W - 1), x1 = min(x0 + 1, W - 1);
  const int y0 = min(max(i + sy, 0), H - 1), y1 = min(y0 + 1, H - 1);
  const long long b = n * H * W;
  o[0] = b + (long long)y0 * W + x0; o[1] = b + (long long)y0 * W + x1;
  o[2] = b + (long long)y1 * W + x0; o[3] = b + (long long)y1 * W + x1;
}

// V1: one float4 per thread, LP=16 lanes per pixel, non-persistent grid, PPT pixels per thread-group sequentially
template <int LP, int QPL, int PPT, bool STORE, int NCORN>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ x, float4* __restrict__ out, int sx, int sy) {
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LP;
  const int lane = threadIdx.x % LP;
  const long long ngrp = (long long)gridDim.x * blockDim.x / LP;
  for (long long pix0 = grp * PPT; pix0 < NPIX; pix0 += ngrp * PPT) {
    float4 v[PPT][QPL][4];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      long long o[4];
      corners(min(pix0 + u, NPIX - 1), sx, sy, o);
#pragma unroll
      for (int q = 0; q < QPL; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) v[u][q][k] = ldnc(x + o[k < NCORN ? k : 0] * C4 + lane + q * LP);
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
      for (int q = 0; q < QPL; ++q) {
        float4 r = comb(v[u][q][0], v[u][q][1], v[u][q][2], v[u][q][3]);
        if (STORE || r.x == 123.456f) { if (pix0 + u < NPIX) __stcs(out + (pix0 + u) * C4 + lane + q * LP, r); }
      }
  }
}

__global__ void k_copy(const float4* __restrict__ x, float4* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = x[i];
}

template <typename F>
float timeit(F f, int iters = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  float4 *x, *out;
  const size_t bytes = (size_t)NPIX * C4 * sizeof(float4);
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&out, bytes));
  CK(cudaMemset(x, 0, bytes));
  const double gb = 2.0 * bytes / 1e9;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  { float ms = timeit([&] { k_copy<<<sms * 32, 256>>>(x, out, (long long)NPIX * C4); }); printf("%-44s %7.3f ms %7.0f GB/s\n", "copy float4 grid-stride", ms, gb / ms * 1e3); }
#define RUN(LP, QPL, PPT, STORE, NC, BLOCKS, NAME)                                                        \
  {                                                                                                        \
    float ms = timeit([&] { k_gather<LP, QPL, PPT, STORE, NC><<<BLOCKS, 256>>>(x, out, 3, 2); });           \
    printf("%-44s %7.3f ms %7.0f GB/s (alg)\n", NAME, ms, gb / ms * 1e3);                                 \
  }
  const int full = (int)(NPIX * 16 / 256);
  RUN(16, 1, 1, true, 4, full, "LP16 Q1 P1 non-persistent")
  RUN(16, 1, 2, true, 4, full / 2, "LP16 Q1 P2 non-persistent")
  RUN(16, 1, 4, true, 4, full / 4, "LP16 Q1 P4 non-persistent")
  RUN(8, 2, 1, true, 4, full / 2, "LP8 Q2 P1 non-persistent")
  RUN(8, 2, 2, true, 4, full / 4, "LP8 Q2 P2 non-persistent")
  RUN(16, 1, 1, true, 4, sms * 8, "LP16 Q1 P1 persistent 8/SM")
  RUN(16, 1, 2, true, 4, sms * 8, "LP16 Q1 P2 persistent 8/SM")
  RUN(16, 1, 4, true, 4, sms * 6, "LP16 Q1 P4 persistent 6/SM")
  RUN(8, 2, 2, true, 4, sms * 6, "LP8 Q2 P2 persistent 6/SM")
  RUN(16, 1, 2, false, 4, full / 2, "LP16 Q1 P2 no-store")
  RUN(16, 1, 2, true, 1, full / 2, "LP16 Q1 P2 one-corner")
  RUN(16, 1, 2, false, 1, full / 2, "LP16 Q1 P2 one-corner no-store")
  RUN(16, 1, 4, false, 1, full / 4, "LP16 Q1 P4 one-corner no-store")
  return 0;
}
