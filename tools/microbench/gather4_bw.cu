// Standalone micro-benchmark (development tool, not product): rows of a [pixels, 64] float32 matrix gathered per
// output pixel, R gathered rows + one 256-byte store per pixel -- the access pattern of the backward's gather kernel
// (R = 16: 8 list rows + 8 corner rows as two roles of 4 + 4 + ...; R = 4: the forward) -- with the rows landing
//   (a) in registers  (ld.global.nc.v4, the shipped kernels), or
//   (b) in shared memory through TMA `cp.async.bulk.tensor.2d.tile::gather4` (UTMALDG.2D.GATHER4; completion on an
//       mbarrier), read back with LDS for the arithmetic.
// Question: do TMA row gathers lift the register wall of the gather kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_bw tools/microbench/gather4_bw.cu && ./gather4_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int W = 512, H = 256, C4 = 16, N = 40;
constexpr long long NPIX = (long long)N * H * W;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 ldnc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// the r-th gathered row of output pixel `pix`: a pseudo-random neighbour within +-6 pixels / +-3 rows (what a smooth
// 8 px flow with 1 px noise produces), clamped into the frame
__device__ __forceinline__ int src_row(long long pix, int r) {
  const int j = (int)(pix % W);
  const long long q = pix / W;
  const int i = (int)(q % H);
  const long long n = q / H;
  unsigned h = (unsigned)pix * 2654435761u + (unsigned)r * 40503u;
  h ^= h >> 15;
  const int dx = (int)(h % 13u) - 6 + 3, dy = (int)((h >> 8) % 7u) - 3 + 2;
  const int x = min(max(j + dx, 0), W - 1), y = min(max(i + dy, 0), H - 1);
  return (int)(n * H * W + (long long)y * W + x);
}

// (a) registers: LP = 8 lanes per pixel, two float4 per lane, R rows in batches of 8 loads in flight per lane
template <int R>
__global__ void __launch_bounds__(256, 4) k_regs(const float4* __restrict__ x, float4* __restrict__ out) {
  const int lane8 = threadIdx.x & 7;
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long ngrp = ((long long)gridDim.x * blockDim.x) >> 3;
  for (long long pix = grp; pix < NPIX; pix += ngrp) {
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += 4) {
      float4 v[4][2];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* row = x + (long long)src_row(pix, r0 + k) * C4 + lane8;
        v[k][0] = ldnc(row);
        v[k][1] = ldnc(row + 8);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a0.x += v[k][0].x; a0.y += v[k][0].y; a0.z += v[k][0].z; a0.w += v[k][0].w;
        a1.x += v[k][1].x; a1.y += v[k][1].y; a1.z += v[k][1].z; a1.w += v[k][1].w;
      }
    }
    __stcs(out + pix * C4 + lane8, a0);
    __stcs(out + pix * C4 + lane8 + 8, a1);
  }
}

// (b) TMA gather4: one warp = a pipeline of SLOTS pixels in flight; lane 0 issues R/4 gather4 operations per pixel
// into the pixel's slot (R x 256 bytes) and arms the slot's mbarrier; the warp then waits for the oldest slot, sums
// its rows (lane l reads 8 bytes of every row) and stores the pixel's row.
template <int R, int SLOTS>
__global__ void __launch_bounds__(256) k_tma(const __grid_constant__ CUtensorMap tm, float2* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slab = reinterpret_cast<float*>(smem) + (size_t)warp * SLOTS * R * 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)8 * SLOTS * R * 256) + warp * SLOTS;
  if (lane == 0)
    for (int s = 0; s < SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  auto issue = [&](long long pix, int s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + s)), "r"(R * 256) : "memory");
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += 4)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
          ::"r"(s32(slab + (size_t)(s * R + r0) * 64)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(bars + s)), "r"(0),
            "r"(src_row(pix, r0)), "r"(src_row(pix, r0 + 1)), "r"(src_row(pix, r0 + 2)), "r"(src_row(pix, r0 + 3))
          : "memory");
  };
  long long head = gw;  // next pixel to issue
  int issued = 0;
  if (lane == 0)
    for (; issued < SLOTS && head < NPIX; ++issued, head += nw) issue(head, issued);
  int s = 0;
  uint32_t phase = 0;
  for (long long pix = gw; pix < NPIX; pix += nw) {
    asm volatile("{\n.reg .pred P1;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 ::"r"(s32(bars + s)), "r"((phase >> s) & 1u) : "memory");
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 v = reinterpret_cast<const float2*>(slab + (size_t)(s * R + r) * 64)[lane];
      acc.x += v.x;
      acc.y += v.y;
    }
    __stcs(out + pix * 32 + lane, acc);
    __syncwarp();  // every lane has read the slot before it is refilled
    phase ^= 1u << s;
    if (lane == 0 && head < NPIX) {
      issue(head, s);
      head += nw;
    }
    s = (s + 1 == SLOTS) ? 0 : s + 1;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

template <int R, int SLOTS>
void run_tma(const CUtensorMap& tm, float4* out, int blocks_per_sm, int sms, double gb) {
  const size_t smem = (size_t)8 * SLOTS * R * 256 + 8 * SLOTS * sizeof(uint64_t);
  CK(cudaFuncSetAttribute(k_tma<R, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float ms = timeit([&] { k_tma<R, SLOTS><<<sms * blocks_per_sm, 256, smem>>>(tm, reinterpret_cast<float2*>(out)); });
  cudaError_t e = cudaGetLastError();
  printf("  TMA gather4   R=%2d slots=%d CTAs/SM=%d (%3zu KB smem/CTA) %8.3f ms %7.0f GB/s%s\n", R, SLOTS, blocks_per_sm, smem >> 10, ms,
         gb / ms * 1e3, e == cudaSuccess ? "" : "  [launch error]");
}

int main() {
  float4 *x, *out;
  const size_t bytes = (size_t)NPIX * C4 * sizeof(float4);
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&out, bytes));
  CK(cudaMemset(x, 0, bytes));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap tm;
  cuuint64_t gdim[2] = {64, (cuuint64_t)NPIX};
  cuuint64_t gstr[1] = {256};
  cuuint32_t box[2] = {64, 1}, estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor map [%lld rows x 64 float32], box {64, 1}: encode rc=%d\n", NPIX, (int)r);
  printf("gathered rows per pixel R, one 256-byte row stored per pixel; GB/s = (R + 1) rows x 256 B x pixels / time (L1-level request rate)\n");
  {
    const double gb = 5.0 * bytes / 1e9;
    float ms = timeit([&] { k_regs<4><<<sms * 4, 256>>>(x, out); });
    printf("  registers     R= 4 (forward pattern)   4 CTAs/SM                %8.3f ms %7.0f GB/s\n", ms, gb / ms * 1e3);
    if (r == CUDA_SUCCESS) {
      run_tma<4, 4>(tm, out, 4, sms, gb);
      run_tma<4, 8>(tm, out, 4, sms, gb);
      run_tma<4, 8>(tm, out, 2, sms, gb);
    }
  }
  {
    const double gb = 17.0 * bytes / 1e9;
    float ms = timeit([&] { k_regs<16><<<sms * 4, 256>>>(x, out); });
    printf("  registers     R=16 (backward pattern)  4 CTAs/SM                %8.3f ms %7.0f GB/s\n", ms, gb / ms * 1e3);
    if (r == CUDA_SUCCESS) {
      run_tma<16, 2>(tm, out, 2, sms, gb);
      run_tma<16, 2>(tm, out, 3, sms, gb);
      run_tma<16, 3>(tm, out, 2, sms, gb);
      run_tma<16, 4>(tm, out, 1, sms, gb);
    }
  }
  return 0;
}
