// warp_bwd.cu -- fused backward of the flow-warp + occlusion blend: grad-input (scatter), grad-flow,
// grad-mask (and grad-other).  Replaces the autograd graph of reference src/utils/ops.py:187-193 +
// src/modules/generator/generator.py:93: mul backward (2 kernels + a C-reduction), ATen
// grid_sampler_2d_backward (zero-fill + 4 atomics per element), add/cat/div backward.
//
// Nothing is saved by the forward except its inputs: coordinates and weights are recomputed here.
//
// Kernels in this file
//   bwd_scatter_kernel    stride-generic, one thread per pixel, channel loop; grad-input by global
//                         red.add (fp32) or, in deterministic mode, by order-independent 64-bit
//                         fixed-point atomics into a workspace that fix2float_kernel converts.
//   absmax_kernel         max |gout * mask| (sets the fixed-point scale; order independent).
//   transpose*_kernel     NCHW <-> channels-last staging copies (C2M_FLAG_STAGE_NHWC).
// launch_bwd() routes a call: channels-last / staged NCHW -> warp_bwd_gather.cu; everything else here.
#include "common.cuh"

namespace c2m {

__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

template <bool DET>
__device__ __forceinline__ void scatter_add(const BwdParams& p, int64_t off, float v, float scale) {
  if (DET) {
    const long long q = __double2ll_rn((double)v * (double)scale);
    atomicAdd(reinterpret_cast<unsigned long long*>(p.acc64 + off), (unsigned long long)q);
  } else {
    red_add(p.gx + off, v);
  }
}

// scale = 2^(61 - count_log2 - exponent(max)), so that count * max * scale < 2^62
__device__ __forceinline__ float fixed_scale(const BwdParams& p) {
  const float mx = __uint_as_float(*p.maxbits);
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &e);
  int k = 61 - p.count_log2 - e;
  k = max(-120, min(120, k));
  return exp2f((float)k);
}

template <bool DET, bool HAS_OTHER>
__global__ void __launch_bounds__(256) bwd_scatter_kernel(const BwdParams p) {
  const Dims& d = p.d;
  const int64_t HW = (int64_t)d.H * d.W;
  const int64_t total = HW * d.N;
  const bool need_x = (p.gflow != nullptr) || (p.gmask != nullptr);
  const float scale = DET ? fixed_scale(p) : 1.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int r = (int)(idx - (int64_t)n * HW);
    const int i = r / d.W, j = r - i * d.W;
    const bool gridmode = (d.flags & C2M_FLAG_COORD_GRID) != 0;
    const float* fl = gridmode ? p.flow + ((int64_t)n * HW + r) * 2 : p.flow + (int64_t)n * 2 * HW + r;
    const float fx = fl[0], fy = fl[gridmode ? 1 : HW];
    const float m = p.mask ? p.mask[(int64_t)n * HW + r] : 1.f;
    Geo g;
    make_geo<true, true>(d, fx, fy, i, j, g);
    const int64_t xbase = (int64_t)(n % d.x_batch) * p.xs[0];
    const int64_t onw = g.y0 * p.xs[2] + g.x0 * p.xs[3], one = g.y0 * p.xs[2] + g.x1 * p.xs[3];
    const int64_t osw = g.y1 * p.xs[2] + g.x0 * p.xs[3], ose = g.y1 * p.xs[2] + g.x1 * p.xs[3];
    const int64_t gb = (int64_t)n * p.gs[0] + i * p.gs[2] + j * p.gs[3];
    float gix = 0.f, giy = 0.f, gm = 0.f;
    for (int c = 0; c < d.C; ++c) {
      const float go = p.gout[gb + c * p.gs[1]];
      const float gg = p.mask ? go * m : go;
      const int64_t xc = xbase + c * p.xs[1];
      if (p.gx) {
        if (g.oknw) scatter_add<DET>(p, xc + onw, g.wnw * gg, scale);
        if (g.okne) scatter_add<DET>(p, xc + one, g.wne * gg, scale);
        if (g.oksw) scatter_add<DET>(p, xc + osw, g.wsw * gg, scale);
        if (g.okse) scatter_add<DET>(p, xc + ose, g.wse * gg, scale);
      }
      if (need_x) {
        const float vnw = g.oknw ? p.x[xc + onw] : 0.f, vne = g.okne ? p.x[xc + one] : 0.f;
        const float vsw = g.oksw ? p.x[xc + osw] : 0.f, vse = g.okse ? p.x[xc + ose] : 0.f;
        // d out / d ix, d out / d iy (ATen grid_sampler_2d_backward)
        gix = fmaf(gg, (vne - vnw) * (1.f - g.ay) + (vse - vsw) * g.ay, gix);
        giy = fmaf(gg, (vsw - vnw) * (1.f - g.ax) + (vse - vne) * g.ax, giy);
        float warped = vnw * g.wnw;
        warped = fmaf(vne, g.wne, warped);
        warped = fmaf(vsw, g.wsw, warped);
        warped = fmaf(vse, g.wse, warped);
        if (HAS_OTHER) warped -= p.other[gb + c * p.gs[1]];
        gm = fmaf(go, warped, gm);
      }
      if (HAS_OTHER && p.gother) p.gother[gb + c * p.gs[1]] = go * (1.f - m);
    }
    if (p.gflow) {
      float* gf = gridmode ? p.gflow + ((int64_t)n * HW + r) * 2 : p.gflow + (int64_t)n * 2 * HW + r;
      gf[0] = gix * g.gmx;
      gf[gridmode ? 1 : HW] = giy * g.gmy;
    }
    if (p.gmask) p.gmask[(int64_t)n * HW + r] = gm;
  }
}

__global__ void __launch_bounds__(256) absmax_kernel(const BwdParams p, unsigned* out_bits) {
  const Dims& d = p.d;
  const int64_t HW = (int64_t)d.H * d.W;
  const int64_t total = HW * d.N;
  float mx = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int r = (int)(idx - (int64_t)n * HW);
    const int i = r / d.W, j = r - i * d.W;
    const float m = p.mask ? fabsf(p.mask[(int64_t)n * HW + r]) : 1.f;
    const int64_t gb = (int64_t)n * p.gs[0] + i * p.gs[2] + j * p.gs[3];
    for (int c = 0; c < d.C; ++c) {
      const float v = fabsf(p.gout[gb + c * p.gs[1]]) * m;
      mx = (v == v) ? fmaxf(mx, v) : __int_as_float(0x7f800000);  // NaN counts as +inf: "non-finite seen"
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out_bits, __float_as_uint(mx));  // non-negative: bits are ordered
}

__global__ void __launch_bounds__(256) fix2float_kernel(const long long* acc, float* gx, int64_t n, BwdParams p) {
  // (NaN when the upstream gradient held a non-finite value: an integer sum cannot carry it)
  const double inv = (double)fixed_inv_scale(__uint_as_float(*p.maxbits), fixed_scale(p));
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x)
    gx[idx] = (float)((double)acc[idx] * inv);
}

// ---------------------------------------------------------------------------------------------
size_t gather_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch);
size_t local_det_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch);
bool gather_supported(const BwdParams& p, Layout lx, Layout lg);
int launch_bwd_gather(const BwdParams& pin, Layout lx, void* workspace, size_t workspace_bytes, cudaStream_t st);

static size_t stage_one(int64_t n, int C, int H, int W) {
  return (((size_t)n * C * H * W * sizeof(float)) + 255) & ~(size_t)255;
}
// channels-last staging copies of gout, x and gx
static size_t stage_bytes(int64_t N, int C, int H, int W, int64_t xb) {
  return stage_one(N, C, H, W) + 2 * stage_one(xb, C, H, W);
}

size_t bwd_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx, int flags) {
  size_t b = 256;
  if (!want_gx) return b;
  const int64_t xb = x_batch > 0 ? x_batch : N;
  if (flags & C2M_FLAG_DETERMINISTIC) {
    // the larger of the two deterministic schemes (the caller does not tell the layout): the fixed-point
    // scatter of the generic kernels, or the channels-last gather with its fixed-point overflow rows
    const size_t a = (size_t)xb * C * H * W * sizeof(long long), l = local_det_workspace_bytes(N, C, H, W, xb);
    b += a > l ? a : l;
    if (flags & C2M_FLAG_STAGE_NHWC) b += stage_bytes(N, C, H, W, xb);
  } else if (!(flags & (C2M_FLAG_BWD_ATOMIC | C2M_FLAG_FORCE_GENERIC | C2M_FLAG_COORD_GRID))) {
    b += gather_workspace_bytes(N, C, H, W, xb);  // contributor lists / candidate lists
    if (flags & C2M_FLAG_STAGE_NHWC) b += stage_bytes(N, C, H, W, xb);
  }
  return b;
}

// ---------------------------------------------------------------------------------------------
// NCHW <-> channels-last staging copies ([n][c][hw] <-> [n][hw][c]) through a 64 x 65 shared-memory tile,
// 128-bit global accesses on both sides (C % 4 == 0, HW % 4 == 0 and 16-byte aligned tensors; transpose_kernel
// below is the scalar fallback).
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) transpose64_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                          int HW) {
  __shared__ float tile[64][65];  // [channel][pixel]
  const int t = threadIdx.x;
  const int64_t img = (int64_t)blockIdx.z * C * HW;
  const int c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const int a = t >> 4, b4 = (t & 15) * 4;  // 16 rows per pass, one float4 column group per thread
  if (TO_NHWC) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // rows = channels, columns = pixels
      const int c = c0 + a + 16 * k, px = p0 + b4;
      if (c < C && px < HW) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(in + img + (int64_t)c * HW + px));
        tile[a + 16 * k][b4] = v.x; tile[a + 16 * k][b4 + 1] = v.y; tile[a + 16 * k][b4 + 2] = v.z; tile[a + 16 * k][b4 + 3] = v.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // rows = pixels, columns = channels
      const int px = p0 + a + 16 * k, c = c0 + b4;
      if (c < C && px < HW)
        *reinterpret_cast<float4*>(out + img + (int64_t)px * C + c) =
            make_float4(tile[b4][a + 16 * k], tile[b4 + 1][a + 16 * k], tile[b4 + 2][a + 16 * k], tile[b4 + 3][a + 16 * k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = p0 + a + 16 * k, c = c0 + b4;
      if (c < C && px < HW) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(in + img + (int64_t)px * C + c));
        tile[b4][a + 16 * k] = v.x; tile[b4 + 1][a + 16 * k] = v.y; tile[b4 + 2][a + 16 * k] = v.z; tile[b4 + 3][a + 16 * k] = v.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + a + 16 * k, px = p0 + b4;
      if (c < C && px < HW)
        __stcs(reinterpret_cast<float4*>(out + img + (int64_t)c * HW + px),
               make_float4(tile[a + 16 * k][b4], tile[a + 16 * k][b4 + 1], tile[a + 16 * k][b4 + 2], tile[a + 16 * k][b4 + 3]));
    }
  }
}

// scalar fallback: 32 x 33 tile, 128-byte rows both ways
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                        int HW) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t img = (int64_t)blockIdx.z * C * HW;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  if (TO_NHWC) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, px = p0 + tx;
      if (c < C && px < HW) tile[ty + 8 * k][tx] = __ldcs(in + img + (int64_t)c * HW + px);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = p0 + ty + 8 * k, c = c0 + tx;
      if (c < C && px < HW) out[img + (int64_t)px * C + c] = tile[tx][ty + 8 * k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = p0 + ty + 8 * k, c = c0 + tx;
      if (c < C && px < HW) tile[ty + 8 * k][tx] = __ldcs(in + img + (int64_t)px * C + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, px = p0 + tx;
      if (c < C && px < HW) __stcs(out + img + (int64_t)c * HW + px, tile[tx][ty + 8 * k]);
    }
  }
}

template <bool TO_NHWC>
static void launch_transpose(const float* in, float* out, int64_t n, int C, int HW, cudaStream_t st) {
  const bool vec = (C % 4) == 0 && (HW % 4) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    const dim3 grid((HW + 63) / 64, (C + 63) / 64, (unsigned)n);
    transpose64_kernel<TO_NHWC><<<grid, 256, 0, st>>>(in, out, C, HW);
  } else {
    const dim3 grid((HW + 31) / 32, (C + 31) / 32, (unsigned)n);
    transpose_kernel<TO_NHWC><<<grid, 256, 0, st>>>(in, out, C, HW);
  }
  count_launch();
}

// C-ABI relayout entry (c2m_relayout): the same staging copy on caller-owned tensors
void launch_relayout(const float* in, float* out, int64_t n, int C, int HW, bool to_nhwc, cudaStream_t st) {
  if (to_nhwc) launch_transpose<true>(in, out, n, C, HW, st);
  else launch_transpose<false>(in, out, n, C, HW, st);
}

static int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

// NCHW-dense tensors through the channels-last kernels: gout (and x, when grad-flow / grad-mask are asked for) are
// copied into channels-last staging buffers, the channels-last backward runs on those, grad-input is copied
// back.  Three streaming passes over a tensor each way cost less than the scattered 4-byte gathers of the
// NCHW kernels.
static int launch_bwd_staged(const BwdParams& pin, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  BwdParams p = pin;
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  char* b = reinterpret_cast<char*>(workspace) + 256;
  float* gout_t = reinterpret_cast<float*>(b);
  b += stage_one(d.N, d.C, d.H, d.W);
  float* x_t = reinterpret_cast<float*>(b);
  b += stage_one(d.x_batch, d.C, d.H, d.W);
  float* gx_t = reinterpret_cast<float*>(b);
  b += stage_one(d.x_batch, d.C, d.H, d.W);
  const size_t used = (size_t)(b - reinterpret_cast<char*>(workspace));
  launch_transpose<true>(p.gout, gout_t, d.N, d.C, HW, st);
  if (p.gflow || p.gmask) launch_transpose<true>(p.x, x_t, d.x_batch, d.C, HW, st);
  float* gx_out = p.gx;
  p.gout = gout_t;
  p.x = x_t;
  if (p.gx) p.gx = gx_t;
  const int64_t nhwc[4] = {(int64_t)d.C * HW, 1, (int64_t)d.W * d.C, d.C};
  for (int k = 0; k < 4; ++k) p.xs[k] = p.gs[k] = nhwc[k];
  const int rc = launch_bwd_gather(p, LAYOUT_NHWC, b, workspace_bytes - used, st);
  if (rc) return rc;
  if (gx_out) launch_transpose<false>(gx_t, gx_out, d.x_batch, d.C, HW, st);
  return C2M_OK;
}

void launch_blend_other_bwd(const BwdParams& p, Layout lg, cudaStream_t st);

int memset_failed() {
  set_error("c2m_warp_blend_bwd: cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
  return C2M_ERR_CUDA;
}

int launch_bwd(const BwdParams& pin, Layout lx, Layout lg, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if ((pin.other || pin.gother) &&
      !(pin.d.flags & (C2M_FLAG_FORCE_GENERIC | C2M_FLAG_BWD_ATOMIC | C2M_FLAG_COORD_GRID | C2M_FLAG_TRUE_DIV | C2M_FLAG_NO_FMA))) {
    // blend operand: the fast kernels run as if `other` were absent (grad-mask = sum_c gout*warp), then one
    // streaming pass writes grad-other = (1-m)*gout and subtracts sum_c gout*other from grad-mask
    BwdParams q = pin;
    q.other = nullptr;
    q.gother = nullptr;
    const int rc = launch_bwd(q, lx, lg, workspace, workspace_bytes, st);
    if (rc) return rc;
    if (pin.other && (pin.gmask || pin.gother)) launch_blend_other_bwd(pin, lg, st);
    return C2M_OK;
  }
  if ((pin.d.flags & C2M_FLAG_STAGE_NHWC) && lx == LAYOUT_NCHW && lg == LAYOUT_NCHW && workspace &&
      workspace_bytes >= bwd_workspace_bytes(pin.d.N, pin.d.C, pin.d.H, pin.d.W, pin.d.x_batch, pin.gx != nullptr,
                                             pin.d.flags) &&
      pin.gx != nullptr) {
    // eligibility of the channels-last kernels, judged on the staging buffers (256-byte aligned)
    BwdParams q = pin;
    q.x = q.gout = reinterpret_cast<const float*>(workspace);
    q.gx = reinterpret_cast<float*>(workspace);
    if (gather_supported(q, LAYOUT_NHWC, LAYOUT_NHWC)) return launch_bwd_staged(pin, workspace, workspace_bytes, st);
  }
  if ((pin.d.flags & C2M_FLAG_PLANNED) &&
      !(lx == LAYOUT_NHWC && pin.gx && !(pin.d.flags & C2M_FLAG_DETERMINISTIC) && gather_supported(pin, lx, lg))) {
    set_error("C2M_FLAG_PLANNED: this call does not take the channels-last float gather the plan was made for");
    return C2M_ERR_INVALID;
  }
  if (gather_supported(pin, lx, lg))
    return launch_bwd_gather(pin, lx, reinterpret_cast<char*>(workspace) + 256,
                             workspace_bytes >= 256 ? workspace_bytes - 256 : 0, st);
  BwdParams p = pin;
  const Dims& d = p.d;
  const int64_t total = (int64_t)d.N * d.H * d.W;
  const int64_t nx = (int64_t)d.x_batch * d.C * d.H * d.W;
  const bool det = (d.flags & C2M_FLAG_DETERMINISTIC) && p.gx;
  const size_t need = bwd_workspace_bytes(d.N, d.C, d.H, d.W, d.x_batch, p.gx != nullptr, d.flags);
  if (det && (workspace == nullptr || workspace_bytes < need)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return C2M_ERR_WORKSPACE;
  }
  if (det) {
    // gx must be dense in memory for the fixed-point accumulator to mirror it
    unsigned* maxbits = reinterpret_cast<unsigned*>(workspace);
    p.acc64 = reinterpret_cast<long long*>(reinterpret_cast<char*>(workspace) + 256);
    p.maxbits = maxbits;
    int cl = 2;  // 4 corners
    int64_t cnt = (int64_t)d.H * d.W * (d.N / d.x_batch);
    while ((1ll << (cl - 2)) < cnt) ++cl;
    p.count_log2 = cl;
    if (cudaMemsetAsync(workspace, 0, 256 + (size_t)nx * sizeof(long long), st) != cudaSuccess) return memset_failed();
    absmax_kernel<<<grid_for(total), 256, 0, st>>>(p, maxbits);
    count_launch();
  } else if (p.gx) {
    if (cudaMemsetAsync(p.gx, 0, (size_t)nx * sizeof(float), st) != cudaSuccess) return memset_failed();
  }
  const int grid = grid_for(total);
  if (det) {
    if (p.other) bwd_scatter_kernel<true, true><<<grid, 256, 0, st>>>(p);
    else bwd_scatter_kernel<true, false><<<grid, 256, 0, st>>>(p);
    count_launch();
    fix2float_kernel<<<grid_for(nx), 256, 0, st>>>(p.acc64, p.gx, nx, p);
    count_launch();
  } else {
    if (p.other) bwd_scatter_kernel<false, true><<<grid, 256, 0, st>>>(p);
    else bwd_scatter_kernel<false, false><<<grid, 256, 0, st>>>(p);
    count_launch();
  }
  return C2M_OK;
}

}  // namespace c2m
