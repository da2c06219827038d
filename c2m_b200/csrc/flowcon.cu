// flowcon.cu -- fused forward-backward flow-consistency loss (second half).
//
// Reference: src/losses/losses.py:115-141 `FlowConsistLoss`:
//     next = mean(mask_fw * |resample(flowback, flow) + flow|)        (masks optional)
//     prev = mean(mask_bw * |resample(flow, flowback) + flowback|)
//     loss = (prev + next) * num_predicted_frames
// on the frame axis folded into the batch (torch.cat(torch.unbind(., 2), 0) of four 5-D tensors).  Each `resample`
// is the C = 2 flavour of the warp (ops.py:187-193: CPU-built grid, H2D copy, five kernels); autograd adds two
// grid_sampler backward kernels with atomics, abs / mul / mean backward and the cat / unbind bookkeeping.
//
// Here: the 5-D tensors are read in place.  Forward = one pass (both terms per pixel) + a one-block finish, sums in
// double in a fixed order.  Backward = one pass + one finish pass: the per-pixel ("direct") parts of the gradient are
// written by the owning thread; the scatter parts -- each flow is also the IMAGE of the other term's warp -- are
// accumulated as 2^-32 fixed point with 64-bit integer atomics (order independent) and folded in by the finish pass.
// The whole loss is therefore bitwise reproducible, unlike the reference's (atomicAdd of floats in ATen).
#include <cstring>

#include "common.cuh"

namespace c2m {

constexpr int kFcBlocks = 2368;
constexpr double kFcFix = 4294967296.0;  // 2^32

struct FcParams {
  Dims d;  // N = B, C = 2
  const float* flow;      // [B,2,T,H,W]
  const float* back;      // [B,2,T,H,W]
  const float* mask_fw;   // [B,1,T,H,W] or NULL
  const float* mask_bw;
  int T;
  int64_t total;          // B*T*H*W
};

struct FcPixel {
  int b, t, r, i, j;
  int64_t o2;  // offset of the x plane in a [B,2,T,H,W] tensor; the y plane is T*HW further
  int64_t o1;  // offset in a [B,1,T,H,W] tensor
};

__device__ __forceinline__ FcPixel fc_decode(const FcParams& p, int64_t idx, int HW) {
  FcPixel q;
  if (p.total <= 0x7fffffff) {  // 32-bit divisions: the 64-bit ones are emulated (~100 instructions each)
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
  q.o2 = (((int64_t)q.b * 2) * p.T + q.t) * HW + q.r;
  q.o1 = ((int64_t)q.b * p.T + q.t) * HW + q.r;
  return q;
}

struct FcTaps {
  float v[2][4];  // [channel][nw, ne, sw, se], zero where the corner is outside the image
};

// sample both planes of `img` (a [B,2,T,H,W] tensor, frame (b, t)) with geometry g
__device__ __forceinline__ void fc_sample(const float* img, const FcPixel& q, int T, int H, int W, const Geo& g, FcTaps& tp,
                                          float& w0, float& w1) {
  const int64_t HW = (int64_t)H * W;
  const float* p0 = img + (((int64_t)q.b * 2) * T + q.t) * HW;
  const float* p1 = p0 + (int64_t)T * HW;
  const int onw = g.y0 * W + g.x0, one = g.y0 * W + g.x1, osw = g.y1 * W + g.x0, ose = g.y1 * W + g.x1;
  // eight loads requested back to back (the corners are clamped into the image: every load is legal), then the
  // out-of-image ones zeroed
  tp.v[0][0] = __ldg(p0 + onw); tp.v[0][1] = __ldg(p0 + one); tp.v[0][2] = __ldg(p0 + osw); tp.v[0][3] = __ldg(p0 + ose);
  tp.v[1][0] = __ldg(p1 + onw); tp.v[1][1] = __ldg(p1 + one); tp.v[1][2] = __ldg(p1 + osw); tp.v[1][3] = __ldg(p1 + ose);
  if (!g.oknw) tp.v[0][0] = tp.v[1][0] = 0.f;
  if (!g.okne) tp.v[0][1] = tp.v[1][1] = 0.f;
  if (!g.oksw) tp.v[0][2] = tp.v[1][2] = 0.f;
  if (!g.okse) tp.v[0][3] = tp.v[1][3] = 0.f;
  w0 = fmaf(tp.v[0][3], g.wse, fmaf(tp.v[0][2], g.wsw, fmaf(tp.v[0][1], g.wne, tp.v[0][0] * g.wnw)));
  w1 = fmaf(tp.v[1][3], g.wse, fmaf(tp.v[1][2], g.wsw, fmaf(tp.v[1][1], g.wne, tp.v[1][0] * g.wnw)));
}

__global__ void __launch_bounds__(256) flowcon_fwd_kernel(const FcParams p, double* __restrict__ partials) {
  __shared__ double s_warp[8];
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  // per-thread sums in float (a thread sees a few dozen terms; the FP64 pipe of this part is 1/64 rate and the
  // double additions were the kernel's bound), per-block and final sums in double in a fixed order
  float acc = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const FcPixel q = fc_decode(p, idx, HW);
    const float fx = __ldg(p.flow + q.o2), fy = __ldg(p.flow + q.o2 + (int64_t)p.T * HW);
    const float bx = __ldg(p.back + q.o2), by = __ldg(p.back + q.o2 + (int64_t)p.T * HW);
    Geo g;
    FcTaps tp;
    float w0, w1;
    make_geo<false>(d, fx, fy, q.i, q.j, g);
    fc_sample(p.back, q, p.T, d.H, d.W, g, tp, w0, w1);
    float a = fabsf(w0 + fx) + fabsf(w1 + fy);
    if (p.mask_fw) {
      const float m = __ldg(p.mask_fw + q.o1);
      a = m * fabsf(w0 + fx) + m * fabsf(w1 + fy);
    }
    make_geo<false>(d, bx, by, q.i, q.j, g);
    fc_sample(p.flow, q, p.T, d.H, d.W, g, tp, w0, w1);
    float b = fabsf(w0 + bx) + fabsf(w1 + by);
    if (p.mask_bw) {
      const float m = __ldg(p.mask_bw + q.o1);
      b = m * fabsf(w0 + bx) + m * fabsf(w1 + by);
    }
    acc += a + b;
  }
  double wacc = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = wacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += s_warp[k];
    partials[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) flowcon_finish_kernel(const double* __restrict__ partials, int n, double numel,
                                                             double scale, float* __restrict__ loss) {
  __shared__ double s[256];
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) a += partials[k];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(s[0] / numel * scale);
}

__device__ __forceinline__ float fc_sign(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : (v == 0.f ? 0.f : v)); }

__device__ __forceinline__ void fc_scatter(long long* acc, const FcPixel& q, int T, int H, int W, const Geo& g, float s0,
                                           float s1) {
  const int64_t HW = (int64_t)H * W;
  long long* a0 = acc + (((int64_t)q.b * 2) * T + q.t) * HW;
  long long* a1 = a0 + (int64_t)T * HW;
  const int off[4] = {g.y0 * W + g.x0, g.y0 * W + g.x1, g.y1 * W + g.x0, g.y1 * W + g.x1};
  const float w[4] = {g.wnw, g.wne, g.wsw, g.wse};
  const bool ok[4] = {g.oknw, g.okne, g.oksw, g.okse};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!ok[k]) continue;
    // (scaling a float by 2^32 is exact, so the conversion needs no double arithmetic -- the FP64 pipe is 1/64 rate)
    if (s0 != 0.f) atomicAdd(reinterpret_cast<unsigned long long*>(a0 + off[k]), (unsigned long long)__float2ll_rn((s0 * w[k]) * 4294967296.f));
    if (s1 != 0.f) atomicAdd(reinterpret_cast<unsigned long long*>(a1 + off[k]), (unsigned long long)__float2ll_rn((s1 * w[k]) * 4294967296.f));
  }
}

// One term of the loss at pixel q: warp `img` by `flo` (fx, fy), d_c = warped_c + flo_c, masked by m.
//   direct gradient to flo: coef*m*(sign(d_c) + coordinate part)   -> dfx, dfy (added)
//   scatter gradient to img: m*sign(d_c)*w_k, fixed point            -> acc_img
//   gradient to the mask: coef * sum_c |d_c|                         -> gm
__device__ __forceinline__ void fc_term_bwd(const FcParams& p, const FcPixel& q, const float* img, float fx, float fy,
                                            float m, float coef, long long* acc_img, float& dfx, float& dfy, float& gm) {
  const Dims& d = p.d;
  Geo g;
  FcTaps tp;
  float w0, w1;
  make_geo<true>(d, fx, fy, q.i, q.j, g);
  fc_sample(img, q, p.T, d.H, d.W, g, tp, w0, w1);
  const float d0 = w0 + fx, d1 = w1 + fy;
  gm = coef * (fabsf(d0) + fabsf(d1));
  const float s0 = fc_sign(d0) * m, s1 = fc_sign(d1) * m;
  const float gix = s0 * ((tp.v[0][1] - tp.v[0][0]) * (1.f - g.ay) + (tp.v[0][3] - tp.v[0][2]) * g.ay) +
                    s1 * ((tp.v[1][1] - tp.v[1][0]) * (1.f - g.ay) + (tp.v[1][3] - tp.v[1][2]) * g.ay);
  const float giy = s0 * ((tp.v[0][2] - tp.v[0][0]) * (1.f - g.ax) + (tp.v[0][3] - tp.v[0][1]) * g.ax) +
                    s1 * ((tp.v[1][2] - tp.v[1][0]) * (1.f - g.ax) + (tp.v[1][3] - tp.v[1][1]) * g.ax);
  dfx += coef * (s0 + gix * g.gmx);
  dfy += coef * (s1 + giy * g.gmy);
  if (acc_img) fc_scatter(acc_img, q, p.T, d.H, d.W, g, s0, s1);
}

__global__ void __launch_bounds__(256) flowcon_bwd_kernel(const FcParams p, const float* __restrict__ gloss, double numel,
                                                          double scale, float* __restrict__ gflow,
                                                          float* __restrict__ gback, float* __restrict__ gmask_fw,
                                                          float* __restrict__ gmask_bw, long long* __restrict__ acc_flow,
                                                          long long* __restrict__ acc_back) {
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const float coef = (float)((double)__ldg(gloss) * scale / numel);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const FcPixel q = fc_decode(p, idx, HW);
    const int64_t oy = q.o2 + (int64_t)p.T * HW;
    const float fx = __ldg(p.flow + q.o2), fy = __ldg(p.flow + oy);
    const float bx = __ldg(p.back + q.o2), by = __ldg(p.back + oy);
    const float mfw = p.mask_fw ? __ldg(p.mask_fw + q.o1) : 1.f, mbw = p.mask_bw ? __ldg(p.mask_bw + q.o1) : 1.f;
    float dfx = 0.f, dfy = 0.f, dbx = 0.f, dby = 0.f, gm1 = 0.f, gm2 = 0.f;
    // next term: image = flowback, flow = flow; prev term: image = flow, flow = flowback
    fc_term_bwd(p, q, p.back, fx, fy, mfw, coef, gback ? acc_back : nullptr, dfx, dfy, gm1);
    fc_term_bwd(p, q, p.flow, bx, by, mbw, coef, gflow ? acc_flow : nullptr, dbx, dby, gm2);
    if (gflow) {
      gflow[q.o2] = dfx;
      gflow[oy] = dfy;
    }
    if (gback) {
      gback[q.o2] = dbx;
      gback[oy] = dby;
    }
    if (gmask_fw) gmask_fw[q.o1] = gm1;
    if (gmask_bw) gmask_bw[q.o1] = gm2;
  }
}

__global__ void __launch_bounds__(256) flowcon_fold_kernel(const float* __restrict__ gloss, double numel, double scale,
                                                           float* __restrict__ gflow, float* __restrict__ gback,
                                                           const long long* __restrict__ acc_flow,
                                                           const long long* __restrict__ acc_back, int64_t n) {
  const double coef = (double)__ldg(gloss) * scale / numel / kFcFix;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    if (gflow) gflow[k] += (float)((double)acc_flow[k] * coef);
    if (gback) gback[k] += (float)((double)acc_back[k] * coef);
  }
}

static int fc_params(FcParams& p, const float* flow, const float* back, const float* mfw, const float* mbw, int64_t B,
                     int T, int H, int W) {
  if (B < 0 || T < 0 || H < 0 || W < 0 || B > 0x7fffffff || (int64_t)H * W > 0x7fffffff) {
    set_error("invalid sizes B=%lld T=%d H=%d W=%d", (long long)B, T, H, W);
    return C2M_ERR_INVALID;
  }
  if ((mfw == nullptr) != (mbw == nullptr)) {
    set_error("mask_fw and mask_bw must be given together (losses.py:123,132)");
    return C2M_ERR_INVALID;
  }
  memset(&p, 0, sizeof(p));
  const int rc = fill_dims(p.d, B, 2, H, W, B, C2M_PAD_BORDER, 0);
  if (rc) return rc;
  p.flow = flow; p.back = back; p.mask_fw = mfw; p.mask_bw = mbw;
  p.T = T;
  p.total = B * T * H * W;
  return C2M_OK;
}

static unsigned fc_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16 < kFcBlocks ? (int64_t)sm_count() * 16 : kFcBlocks;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

static size_t fc_acc_bytes(int64_t B, int T, int H, int W) {
  return (((size_t)B * 2 * T * H * W * sizeof(long long)) + 255) & ~(size_t)255;
}

}  // namespace c2m

using namespace c2m;

extern "C" {

size_t c2m_flow_consistency_workspace_bytes(int64_t B, int T, int H, int W) {
  if (B < 0 || T < 0 || H < 0 || W < 0) return 0;
  return (size_t)kFcBlocks * sizeof(double) + 2 * fc_acc_bytes(B, T, H, W);
}

int c2m_flow_consistency_fwd(const float* flow, const float* flowback, const float* mask_fw, const float* mask_bw,
                             float* loss, int64_t B, int T, int H, int W, float scale, void* workspace,
                             size_t workspace_bytes, void* cuda_stream) {
  FcParams p;
  int rc = fc_params(p, flow, flowback, mask_fw, mask_bw, B, T, H, W);
  if (rc) return rc;
  if (!loss) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if (!workspace || workspace_bytes < (size_t)kFcBlocks * sizeof(double)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, (size_t)kFcBlocks * sizeof(double));
    return C2M_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  double* partials = reinterpret_cast<double*>(workspace);
  const double numel = (double)p.total * 2.0;
  int nb = 0;
  if (numel > 0) {
    if (!flow || !flowback) {
      set_error("null pointer argument");
      return C2M_ERR_INVALID;
    }
    nb = (int)fc_grid(p.total);
    flowcon_fwd_kernel<<<nb, 256, 0, st>>>(p, partials);
    count_launch();
  }
  flowcon_finish_kernel<<<1, 256, 0, st>>>(partials, nb, numel, (double)scale, loss);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_flow_consistency_fwd: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

int c2m_flow_consistency_bwd(const float* flow, const float* flowback, const float* mask_fw, const float* mask_bw,
                             const float* gloss, float* gflow, float* gflowback, float* gmask_fw, float* gmask_bw,
                             int64_t B, int T, int H, int W, float scale, void* workspace, size_t workspace_bytes,
                             void* cuda_stream) {
  FcParams p;
  int rc = fc_params(p, flow, flowback, mask_fw, mask_bw, B, T, H, W);
  if (rc) return rc;
  if (p.total == 0 || (!gflow && !gflowback && !gmask_fw && !gmask_bw)) return C2M_OK;
  if (!flow || !flowback || !gloss) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if ((gmask_fw && !mask_fw) || (gmask_bw && !mask_bw)) {
    set_error("mask gradient requested without a mask");
    return C2M_ERR_INVALID;
  }
  const size_t need = c2m_flow_consistency_workspace_bytes(B, T, H, W);
  if (!workspace || workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return C2M_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  char* b = reinterpret_cast<char*>(workspace) + (size_t)kFcBlocks * sizeof(double);
  long long* acc_flow = reinterpret_cast<long long*>(b);
  long long* acc_back = reinterpret_cast<long long*>(b + fc_acc_bytes(B, T, H, W));
  if (gflow || gflowback) {
    if (cudaMemsetAsync(b, 0, 2 * fc_acc_bytes(B, T, H, W), st) != cudaSuccess) {
      set_error("c2m_flow_consistency_bwd: cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
      return C2M_ERR_CUDA;
    }
  }
  const double numel = (double)p.total * 2.0;
  flowcon_bwd_kernel<<<fc_grid(p.total), 256, 0, st>>>(p, gloss, numel, (double)scale, gflow, gflowback, gmask_fw,
                                                       gmask_bw, acc_flow, acc_back);
  count_launch();
  if (gflow || gflowback) {
    flowcon_fold_kernel<<<fc_grid(p.total * 2), 256, 0, st>>>(gloss, numel, (double)scale, gflow, gflowback, acc_flow,
                                                             acc_back, p.total * 2);
    count_launch();
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_flow_consistency_bwd: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

}  // extern "C"
