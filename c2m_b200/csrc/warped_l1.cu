// warped_l1.cu -- fused warped-frame L1 loss.
//
//   loss = mean over (b, c, t, i, j) of | resample(source, flows[:, :, t])[b, c, i, j] - targets[b, c, t, i, j] |
//
// Reference: src/losses/losses.py:219-222 -- T calls of utils.resample (ops.py:187-193) on the C = 3 source
// frame, torch.cat of the T warped frames, then L1MaskedLoss without a mask (losses.py:184-189 -> F.l1_loss).
// Only the flows carry a gradient there (the frames are data).  The fused form reads the source frame, the 5-D
// flow and target tensors in place (no per-frame slices, no warped clip in memory): one pass forward, one pass
// backward.  The sum is deterministic: per-block partial sums in double, added in block order by one block.
#include "common.cuh"

namespace c2m {

constexpr int kL1Blocks = 2368;  // upper bound of the grid (16 per SM); one partial sum each

struct L1Params {
  Dims d;  // N = B (frames of the source), C, H, W; border padding
  const float* src;    // [B, C, H, W]
  const float* flows;  // [B, 2, T, H, W]
  const float* tgt;    // [B, C, T, H, W]
  int T;
  int64_t total;       // B * T * H * W pixels
};

struct L1Pixel {
  int b, t, r, i, j;
};

__device__ __forceinline__ L1Pixel l1_decode(const L1Params& p, int64_t idx) {
  const int HW = p.d.H * p.d.W;
  L1Pixel q;
  if (p.total <= 0x7fffffff) {  // 32-bit divisions: the 64-bit ones are emulated and would dominate the kernel
    const unsigned u = (unsigned)idx, f = u / (unsigned)HW;
    q.r = (int)(u - f * (unsigned)HW);
    q.b = (int)(f / (unsigned)p.T);
    q.t = (int)(f - (unsigned)q.b * (unsigned)p.T);
  } else {
    const int64_t f = idx / HW;
    q.r = (int)(idx - f * HW);
    q.b = (int)(f / p.T);
    q.t = (int)(f - (int64_t)q.b * p.T);
  }
  q.i = (int)((unsigned)q.r / (unsigned)p.d.W);
  q.j = q.r - q.i * p.d.W;
  return q;
}

__device__ __forceinline__ void l1_fetch(const L1Params& p, int64_t idx, int HW, L1Pixel& q, float& fx, float& fy) {
  q = l1_decode(p, idx);
  const float* fl = p.flows + ((int64_t)q.b * 2 * p.T + q.t) * HW + q.r;
  fx = __ldg(fl);
  fy = __ldg(fl + (int64_t)p.T * HW);
}

// bilinear sample of channel plane `xc`, in the accumulation order of the warp kernels (bit-equal to ATen's)
__device__ __forceinline__ float l1_sample(const float* xc, const Geo& g, int W, float& vnw, float& vne, float& vsw,
                                           float& vse) {
  vnw = g.oknw ? __ldg(xc + g.y0 * W + g.x0) : 0.f;
  vne = g.okne ? __ldg(xc + g.y0 * W + g.x1) : 0.f;
  vsw = g.oksw ? __ldg(xc + g.y1 * W + g.x0) : 0.f;
  vse = g.okse ? __ldg(xc + g.y1 * W + g.x1) : 0.f;
  return fmaf(vse, g.wse, fmaf(vsw, g.wsw, fmaf(vne, g.wne, vnw * g.wnw)));
}

// The four corner values of CT channel planes, requested back to back before the first use (the corners are clamped
// into the image, so every load is legal; out-of-image ones are zeroed afterwards): one exposed memory round trip per
// pixel instead of one per channel.
template <int CT>
__device__ __forceinline__ void l1_gather(const float* xc, const Geo& g, int W, int HW, float (&v)[CT][4]) {
  const int onw = g.y0 * W + g.x0, one = g.y0 * W + g.x1, osw = g.y1 * W + g.x0, ose = g.y1 * W + g.x1;
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    v[c][0] = __ldg(xc + c * HW + onw);
    v[c][1] = __ldg(xc + c * HW + one);
    v[c][2] = __ldg(xc + c * HW + osw);
    v[c][3] = __ldg(xc + c * HW + ose);
  }
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    if (!g.oknw) v[c][0] = 0.f;
    if (!g.okne) v[c][1] = 0.f;
    if (!g.oksw) v[c][2] = 0.f;
    if (!g.okse) v[c][3] = 0.f;
  }
}

// CT: compile-time channel count (3: the frames of losses.py:219-222), 0 = run-time channel loop
template <int CT>
__global__ void __launch_bounds__(256, 4) warped_l1_fwd_kernel(const L1Params p, double* __restrict__ partials) {
  __shared__ float s_warp[8];
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the next pixel's flow is fetched while this one is processed: one exposed memory round trip per pixel
  L1Pixel qn;
  float nfx = 0.f, nfy = 0.f;
  if (idx < p.total) l1_fetch(p, idx, HW, qn, nfx, nfy);
  for (; idx < p.total; idx += stride) {
    const L1Pixel q = qn;
    const float fx = nfx, fy = nfy;
    if (idx + stride < p.total) l1_fetch(p, idx + stride, HW, qn, nfx, nfy);
    Geo g;
    make_geo<false>(d, fx, fy, q.i, q.j, g);
    const float* xc = p.src + (int64_t)q.b * d.C * HW;
    const float* tc = p.tgt + ((int64_t)q.b * d.C * p.T + q.t) * HW + q.r;
    if (CT > 0) {
      float v[CT > 0 ? CT : 1][4], tg[CT > 0 ? CT : 1];
      l1_gather<(CT > 0 ? CT : 1)>(xc, g, d.W, HW, v);
#pragma unroll
      for (int c = 0; c < CT; ++c) tg[c] = __ldg(tc + (int64_t)c * p.T * HW);
#pragma unroll
      for (int c = 0; c < CT; ++c)
        acc += fabsf(fmaf(v[c][3], g.wse, fmaf(v[c][2], g.wsw, fmaf(v[c][1], g.wne, v[c][0] * g.wnw))) - tg[c]);
    } else {
      for (int c = 0; c < d.C; ++c) {
        float a, b2, c2, e;
        const float w = l1_sample(xc, g, d.W, a, b2, c2, e);
        acc += fabsf(w - __ldg(tc));
        xc += HW;
        tc += (int64_t)p.T * HW;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (double)s_warp[k];
    partials[blockIdx.x] = s;
  }
}

// one block: the partial sums in a fixed order, divided by the element count (0 elements -> NaN, as torch's mean)
__global__ void __launch_bounds__(256) warped_l1_finish_kernel(const double* __restrict__ partials, int n, double numel,
                                                               float* __restrict__ loss) {
  __shared__ double s[256];
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) a += partials[k];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(s[0] / numel);
}

// d loss / d flows (and, optionally, d loss / d targets): g = gloss / numel, s_c = sign(warped_c - target_c)
//   gflow_x = g * sum_c s_c * d warped_c / d ix * d ix / d flow_x     (the coordinate algebra of make_geo<true>)
template <int CT>
__global__ void __launch_bounds__(256, 3) warped_l1_bwd_kernel(const L1Params p, const float* __restrict__ gloss,
                                                            double numel, float* __restrict__ gflows,
                                                            float* __restrict__ gtargets) {
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const float gs = (float)((double)__ldg(gloss) / numel);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  L1Pixel qn;
  float nfx = 0.f, nfy = 0.f;
  if (idx < p.total) l1_fetch(p, idx, HW, qn, nfx, nfy);
  for (; idx < p.total; idx += stride) {
    const L1Pixel q = qn;
    const float fx = nfx, fy = nfy;
    if (idx + stride < p.total) l1_fetch(p, idx + stride, HW, qn, nfx, nfy);
    const int64_t fo = ((int64_t)q.b * 2 * p.T + q.t) * HW + q.r;
    Geo g;
    make_geo<true>(d, fx, fy, q.i, q.j, g);
    const float* xc = p.src + (int64_t)q.b * d.C * HW;
    int64_t to = ((int64_t)q.b * d.C * p.T + q.t) * HW + q.r;
    float gix = 0.f, giy = 0.f;
    // one channel's share of the gradients (sg = sign(warped - target) * gloss / numel; NaN stays NaN)
    auto term = [&](float vnw, float vne, float vsw, float vse, float tgt, int64_t at) {
      const float diff = fmaf(vse, g.wse, fmaf(vsw, g.wsw, fmaf(vne, g.wne, vnw * g.wnw))) - tgt;
      const float sg = diff > 0.f ? gs : (diff < 0.f ? -gs : (diff == 0.f ? 0.f : diff));
      gix = fmaf(sg, (vne - vnw) * (1.f - g.ay) + (vse - vsw) * g.ay, gix);
      giy = fmaf(sg, (vsw - vnw) * (1.f - g.ax) + (vse - vne) * g.ax, giy);
      if (gtargets) gtargets[at] = -sg;
    };
    if (CT > 0) {
      float v[CT > 0 ? CT : 1][4], tg[CT > 0 ? CT : 1];
      l1_gather<(CT > 0 ? CT : 1)>(xc, g, d.W, HW, v);
#pragma unroll
      for (int c = 0; c < CT; ++c) tg[c] = __ldg(p.tgt + to + (int64_t)c * p.T * HW);
#pragma unroll
      for (int c = 0; c < CT; ++c) term(v[c][0], v[c][1], v[c][2], v[c][3], tg[c], to + (int64_t)c * p.T * HW);
    } else {
      for (int c = 0; c < d.C; ++c) {
        float vnw, vne, vsw, vse;
        (void)l1_sample(xc, g, d.W, vnw, vne, vsw, vse);
        term(vnw, vne, vsw, vse, __ldg(p.tgt + to), to);
        xc += HW;
        to += (int64_t)p.T * HW;
      }
    }
    if (gflows) {
      gflows[fo] = gix * g.gmx;
      gflows[fo + (int64_t)p.T * HW] = giy * g.gmy;
    }
  }
}

static int l1_params(L1Params& p, const float* source, const float* flows, const float* targets, int64_t B, int C,
                     int T, int H, int W) {
  if (B < 0 || C < 0 || T < 0 || H < 0 || W < 0 || B > 0x7fffffff || (int64_t)H * W > 0x7fffffff) {
    set_error("invalid sizes B=%lld C=%d T=%d H=%d W=%d", (long long)B, C, T, H, W);
    return C2M_ERR_INVALID;
  }
  memset(&p, 0, sizeof(p));
  const int rc = fill_dims(p.d, B, C, H, W, B, C2M_PAD_BORDER, 0);
  if (rc) return rc;
  p.src = source;
  p.flows = flows;
  p.tgt = targets;
  p.T = T;
  p.total = B * T * H * W;
  return C2M_OK;
}

static unsigned l1_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16 < kL1Blocks ? (int64_t)sm_count() * 16 : kL1Blocks;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace c2m

using namespace c2m;

extern "C" {

size_t c2m_warped_l1_workspace_bytes(void) { return (size_t)kL1Blocks * sizeof(double); }

int c2m_warped_l1_fwd(const float* source, const float* flows, const float* targets, float* loss, int64_t B, int C,
                      int T, int H, int W, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  L1Params p;
  int rc = l1_params(p, source, flows, targets, B, C, T, H, W);
  if (rc) return rc;
  if (!loss) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if (!workspace || workspace_bytes < c2m_warped_l1_workspace_bytes()) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, c2m_warped_l1_workspace_bytes());
    return C2M_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  double* partials = reinterpret_cast<double*>(workspace);
  const double numel = (double)p.total * (double)C;
  int nb = 0;
  if (numel > 0) {
    if (!source || !flows || !targets) {
      set_error("null pointer argument");
      return C2M_ERR_INVALID;
    }
    nb = (int)l1_grid(p.total);
    if (C == 3) warped_l1_fwd_kernel<3><<<nb, 256, 0, st>>>(p, partials);
    else warped_l1_fwd_kernel<0><<<nb, 256, 0, st>>>(p, partials);
    count_launch();
  }
  warped_l1_finish_kernel<<<1, 256, 0, st>>>(partials, nb, numel, loss);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_warped_l1_fwd: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

int c2m_warped_l1_bwd(const float* source, const float* flows, const float* targets, const float* gloss,
                      float* gflows, float* gtargets, int64_t B, int C, int T, int H, int W, void* cuda_stream) {
  L1Params p;
  int rc = l1_params(p, source, flows, targets, B, C, T, H, W);
  if (rc) return rc;
  if (p.total == 0 || (!gflows && !gtargets)) return C2M_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (C == 0) {  // no channels: the flows' gradient is zero (the loss itself is NaN)
    if (gflows && cudaMemsetAsync(gflows, 0, (size_t)p.total * 2 * sizeof(float), st) != cudaSuccess) {
      set_error("c2m_warped_l1_bwd: cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
      return C2M_ERR_CUDA;
    }
    return C2M_OK;
  }
  if (!source || !flows || !targets || !gloss) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if (C == 3)
    warped_l1_bwd_kernel<3><<<l1_grid(p.total), 256, 0, st>>>(p, gloss, (double)p.total * (double)C, gflows, gtargets);
  else
    warped_l1_bwd_kernel<0><<<l1_grid(p.total), 256, 0, st>>>(p, gloss, (double)p.total * (double)C, gflows, gtargets);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_warped_l1_bwd: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

}  // extern "C"
