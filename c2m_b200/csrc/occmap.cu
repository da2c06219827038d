// occmap.cu -- forward-splat occlusion map, the producer of the warp's mask
// (reference src/utils/ops.py:205-275: get_corresponding_map / mesh_grid / get_occlusion_map; called per frame at
// src/modules/motion_estimator/dense_motion.py:148,151).
//
// Every pixel (i, j) moves to (j + fx, i + fy) in plain pixel coordinates and splats its bilinear weights onto the
// four surrounding pixels; corners outside the image are dropped; the occlusion map is the sum clamped to [0, 1].
// The reference does the sum with scatter_add_ (float atomics on CUDA: order-dependent rounding).  Here the weights
// (all in [0, 1]) are accumulated as 2^-32 fixed point in 64-bit integers -- the same scatter as the deterministic
// grad-input, with C = 1 -- so the result is independent of the order and bitwise reproducible.
#include "common.cuh"

namespace c2m {

constexpr float kOccScale = 4294967296.f;        // 2^32
constexpr double kOccInv = 1.0 / 4294967296.0;

template <bool COORDS>
__global__ void __launch_bounds__(256) occmap_scatter_kernel(const float* __restrict__ in, unsigned long long* acc,
                                                             int64_t N, int H, int W) {
  const int64_t HW = (int64_t)H * W, total = HW * N;
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / HW;
    const int r = (int)(idx - n * HW);
    const int i = r / W, j = r - i * W;
    float x = __ldg(in + n * 2 * HW + r), y = __ldg(in + n * 2 * HW + HW + r);
    if (!COORDS) {  // ops.py:270-273: mesh grid + flow, one fp32 add
      x = __fadd_rn((float)j, x);
      y = __fadd_rn((float)i, y);
    }
    if (!(x == x) || !(y == y)) continue;  // NaN: the reference's index cast is undefined there; drop the pixel
    // ops.py:216-223
    const float x1 = floorf(x), y1 = floorf(y);
    const float xf = fminf(fmaxf(x1, 0.f), wm), yf = fminf(fmaxf(y1, 0.f), hm);
    const float x0 = x1 + 1.f, y0 = y1 + 1.f;
    const float xc = fminf(fmaxf(x0, 0.f), wm), yc = fminf(fmaxf(y0, 0.f), hm);
    const bool xc_out = x0 != xc, yc_out = y0 != yc, xf_out = x1 != xf, yf_out = y1 != yf;
    const float wxc = 1.f - fabsf(x - xc), wxf = 1.f - fabsf(x - xf);
    const float wyc = 1.f - fabsf(y - yc), wyf = 1.f - fabsf(y - yf);
    unsigned long long* a = acc + n * HW;
    // ops.py:237-247 (order: ceil/ceil, ceil/floor, floor/ceil, floor/floor)
    if (!(xc_out | yc_out)) atomicAdd(a + (int)xc + (int)yc * W, (unsigned long long)__float2ll_rn((wxc * wyc) * kOccScale));
    if (!(xc_out | yf_out)) atomicAdd(a + (int)xc + (int)yf * W, (unsigned long long)__float2ll_rn((wxc * wyf) * kOccScale));
    if (!(xf_out | yc_out)) atomicAdd(a + (int)xf + (int)yc * W, (unsigned long long)__float2ll_rn((wxf * wyc) * kOccScale));
    if (!(xf_out | yf_out)) atomicAdd(a + (int)xf + (int)yf * W, (unsigned long long)__float2ll_rn((wxf * wyf) * kOccScale));
  }
}

template <bool CLAMP>
__global__ void __launch_bounds__(256) occmap_finish_kernel(const long long* __restrict__ acc, float* __restrict__ out,
                                                            int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    float v = (float)((double)acc[idx] * kOccInv);
    if (CLAMP) v = fminf(fmaxf(v, 0.f), 1.f);  // ops.py:275
    out[idx] = v;
  }
}

}  // namespace c2m

using namespace c2m;

extern "C" {

size_t c2m_occlusion_map_workspace_bytes(int64_t N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)N * H * W * sizeof(long long);
}

int c2m_occlusion_map(const float* in, float* out, int64_t N, int H, int W, int flags, void* workspace,
                      size_t workspace_bytes, void* cuda_stream) {
  if (N < 0 || H < 0 || W < 0 || N > 0x7fffffff) {
    set_error("invalid sizes N=%lld H=%d W=%d", (long long)N, H, W);
    return C2M_ERR_INVALID;
  }
  if (N == 0 || H == 0 || W == 0) return C2M_OK;
  if (!in || !out) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if ((int64_t)H * W >= (1ll << 24)) {  // the reference encodes the index in fp32 (ops.py:234-237)
    set_error("image too large for the occlusion map (H*W must be < 2^24)");
    return C2M_ERR_INVALID;
  }
  const size_t need = c2m_occlusion_map_workspace_bytes(N, H, W);
  if (!workspace || workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return C2M_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (cudaMemsetAsync(workspace, 0, need, st) != cudaSuccess) {
    set_error("c2m_occlusion_map: cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
    return C2M_ERR_CUDA;
  }
  const int64_t total = N * H * W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(workspace);
  if (flags & C2M_OCC_COORDS) occmap_scatter_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(in, acc, N, H, W);
  else occmap_scatter_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(in, acc, N, H, W);
  if (flags & C2M_OCC_NO_CLAMP)
    occmap_finish_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(workspace), out, total);
  else
    occmap_finish_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(workspace), out, total);
  count_launch(2);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_occlusion_map: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

}  // extern "C"
