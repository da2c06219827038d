// warp_aux.cu -- the small passes around the fused warp's backward:
//
//   resize_fwd_kernel   flow / mask resized to the feature size (the forward kernels do this on the fly,
//                       common.cuh:fetch_flow_mask; the backward materialises them once in its workspace so that the
//                       TMA-staged kernels run unchanged).  Reference: generator.py:84-85,91-92 (F.interpolate
//                       bilinear, align_corners=False), utils.py:346-354 (align_corners=True + value rescale),
//                       motion_autoencoder.py:120-124.
//   resize_bwd_kernel   back-propagates grad-flow / grad-mask from the feature size to the size the caller holds them
//                       at (autograd of the above: ATen upsample_bilinear2d_backward + the in-place divides), as a
//                       GATHER over source pixels -- no zero-fill, no atomics, one fixed summation order.
//   blend_other_bwd_kernel   the `other` operand of out = m*warp + (1-m)*other (north-star blend):
//                       grad-other = (1-m)*gout and grad-mask -= sum_c gout*other, after the main backward.
#include <cstring>

#include "common.cuh"

namespace c2m {

int fill_resize(Dims& d, const c2m_resize* rs) {
  Resize& r = d.rs;
  memset(&r, 0, sizeof(r));
  r.fmulx = r.fmuly = 1.f;
  if (!rs) return C2M_OK;
  const int fh = rs->flow_h > 0 ? rs->flow_h : d.H, fw = rs->flow_w > 0 ? rs->flow_w : d.W;
  const int mh = rs->mask_h > 0 ? rs->mask_h : d.H, mw = rs->mask_w > 0 ? rs->mask_w : d.W;
  if (rs->flow_mode != C2M_RESIZE_HALF_PIXEL && rs->flow_mode != C2M_RESIZE_CORNERS_RESCALE) {
    set_error("invalid flow_mode %d", rs->flow_mode);
    return C2M_ERR_INVALID;
  }
  if (d.flags & C2M_FLAG_COORD_GRID) {
    if (fh != d.H || fw != d.W || mh != d.H || mw != d.W) {
      set_error("a normalised sampling grid cannot be resized");
      return C2M_ERR_INVALID;
    }
    return C2M_OK;
  }
  r.fh = fh; r.fw = fw; r.mh = mh; r.mw = mw;
  r.f_align = rs->flow_mode == C2M_RESIZE_CORNERS_RESCALE;
  // utils.py:346-354 always interpolates (identity when the sizes match, bit for bit) -- only a real size change
  // needs work here
  if (fh != d.H || fw != d.W) r.on |= 1;
  if (mh != d.H || mw != d.W) r.on |= 2;
  if (rs->fold_t < 0 || (rs->fold_t > 0 && d.N % rs->fold_t != 0)) {
    set_error("fold_t=%d must divide N=%d", rs->fold_t, d.N);
    return C2M_ERR_INVALID;
  }
  if (rs->fold_t > 0) {  // 5-D clips: every plane is addressed through the fold, resized or not
    r.fold_t = rs->fold_t;
    r.fold_b = d.N / rs->fold_t;
    r.on |= 3;
  }
  // ATen area_pixel_compute_scale<float>: align_corners ? (in-1)/(out-1) (0 when out == 1) : in/out
  if (r.f_align) {
    r.fsy = d.H > 1 ? (float)(fh - 1) / (float)(d.H - 1) : 0.f;
    r.fsx = d.W > 1 ? (float)(fw - 1) / (float)(d.W - 1) : 0.f;
    // out[:, 0] /= w / float(new_w): a python double, cast to float32, applied as a reciprocal multiply on CUDA
    r.fmulx = 1.0f / (float)((double)fw / (double)d.W);
    r.fmuly = 1.0f / (float)((double)fh / (double)d.H);
  } else {
    r.fsy = (float)fh / (float)d.H;
    r.fsx = (float)fw / (float)d.W;
  }
  r.msy = (float)mh / (float)d.H;
  r.msx = (float)mw / (float)d.W;
  return C2M_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_fwd_kernel(const Dims d, const float* __restrict__ flow,
                                                         const float* __restrict__ mask, float* __restrict__ flow_out,
                                                         float* __restrict__ mask_out) {
  const int HW = d.H * d.W;
  const int64_t total = (int64_t)HW * d.N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / HW);
    const int r = (int)(idx - (int64_t)n * HW);
    const int i = r / d.W, j = r - i * d.W;
    float fx, fy, m;
    fetch_flow_mask(d, flow, mask, n, i, j, fx, fy, m);
    flow_out[(int64_t)n * 2 * HW + r] = fx;
    flow_out[(int64_t)n * 2 * HW + HW + r] = fy;
    if (mask_out) mask_out[idx] = m;
  }
}

// Destination pixels whose taps can touch source index s along one axis: a conservative range from the inverse of
// the source-index map; the exact test (same float arithmetic as the forward) happens in the loop.
__device__ __forceinline__ void dest_range(float scale, int s, int out_size, bool align, int& lo, int& hi) {
  if (scale <= 0.f) {  // out_size == 1 with align_corners: every destination reads source 0
    lo = 0;
    hi = out_size - 1;
    return;
  }
  const float inv = 1.f / scale;
  float a, b;
  if (align) {
    a = ((float)s - 1.f) * inv;
    b = ((float)s + 1.f) * inv;
  } else {
    a = ((float)s - 0.5f) * inv - 0.5f;
    b = ((float)s + 1.5f) * inv - 0.5f;
  }
  lo = max(0, (int)floorf(a) - 1);
  hi = min(out_size - 1, (int)ceilf(b) + 1);
}

// weight with which destination index `dst` reads source index `s` along one axis
__device__ __forceinline__ float tap_weight(float scale, int dst, int in_size, bool align, int s) {
  int i0, ip;
  float l0, l1;
  resize_taps(scale, dst, in_size, align, i0, ip, l0, l1);
  float w = 0.f;
  if (i0 == s) w += l0;
  if (i0 + ip == s) w += l1;
  return w;
}

// One thread per SOURCE pixel of one plane kind: grad_src[s] = sum over destinations of wy * wx * grad_dst.
// planes: which == 0 -> the two flow planes (value rescale folded in), which == 1 -> the mask plane.
__global__ void __launch_bounds__(256) resize_bwd_kernel(const Dims d, const float* __restrict__ gflow_small,
                                                         const float* __restrict__ gmask_small,
                                                         float* __restrict__ gflow_src, float* __restrict__ gmask_src) {
  const Resize& rs = d.rs;
  const bool do_flow = gflow_src != nullptr && (rs.on & 1), do_mask = gmask_src != nullptr && (rs.on & 2);
  const int64_t nf = do_flow ? (int64_t)d.N * rs.fh * rs.fw : 0;
  const int64_t nm = do_mask ? (int64_t)d.N * rs.mh * rs.mw : 0;
  const int HW = d.H * d.W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nf + nm;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const bool is_flow = idx < nf;
    const int64_t k = is_flow ? idx : idx - nf;
    const int Hs = is_flow ? rs.fh : rs.mh, Ws = is_flow ? rs.fw : rs.mw;
    const float sy = is_flow ? rs.fsy : rs.msy, sx = is_flow ? rs.fsx : rs.msx;
    const bool align = is_flow && rs.f_align;
    const int n = (int)(k / ((int64_t)Hs * Ws));
    const int r = (int)(k - (int64_t)n * Hs * Ws);
    const int ys = r / Ws, xs = r - ys * Ws;
    int ylo, yhi, xlo, xhi;
    dest_range(sy, ys, d.H, align, ylo, yhi);
    dest_range(sx, xs, d.W, align, xlo, xhi);
    float a0 = 0.f, a1 = 0.f;
    for (int i = ylo; i <= yhi; ++i) {
      const float wy = tap_weight(sy, i, Hs, align, ys);
      if (wy == 0.f) continue;
      for (int j = xlo; j <= xhi; ++j) {
        const float wx = tap_weight(sx, j, Ws, align, xs);
        if (wx == 0.f) continue;
        const float w = wy * wx;
        if (is_flow) {
          const float* g = gflow_small + (int64_t)n * 2 * HW + i * d.W + j;
          a0 = fmaf(w, __ldg(g), a0);
          a1 = fmaf(w, __ldg(g + HW), a1);
        } else {
          a0 = fmaf(w, __ldg(gmask_small + (int64_t)n * HW + i * d.W + j), a0);
        }
      }
    }
    if (is_flow) {
      gflow_src[plane_offset(rs, n, 0, 2, (int64_t)Hs * Ws) + r] = a0 * rs.fmulx;
      gflow_src[plane_offset(rs, n, 1, 2, (int64_t)Hs * Ws) + r] = a1 * rs.fmuly;
    } else {
      gmask_src[plane_offset(rs, n, 0, 1, (int64_t)Hs * Ws) + r] = a0;
    }
  }
}

static int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

void launch_resize_fwd(const Dims& d, const float* flow_src, const float* mask_src, float* flow_out, float* mask_out,
                       cudaStream_t st) {
  resize_fwd_kernel<<<grid_for((int64_t)d.N * d.H * d.W), 256, 0, st>>>(d, flow_src, mask_src, flow_out,
                                                                         mask_src ? mask_out : nullptr);
  count_launch();
}

void launch_resize_bwd(const Dims& d, const float* gflow_small, const float* gmask_small, float* gflow_src,
                       float* gmask_src, cudaStream_t st) {
  int64_t total = 0;
  if (gflow_src && (d.rs.on & 1)) total += (int64_t)d.N * d.rs.fh * d.rs.fw;
  if (gmask_src && (d.rs.on & 2)) total += (int64_t)d.N * d.rs.mh * d.rs.mw;
  if (total == 0) return;
  resize_bwd_kernel<<<grid_for(total), 256, 0, st>>>(d, gflow_small, gmask_small, gflow_src, gmask_src);
  count_launch();
}

// ---------------------------------------------------------------------------------------------
// Blend operand.  out = m*warp(x) + (1-m)*other  =>  grad-other = (1-m)*gout,  grad-mask = sum_c gout*(warp - other).
// The main backward runs as if `other` were absent (grad-mask = sum_c gout*warp); this pass writes grad-other and
// subtracts sum_c gout*other from grad-mask.  VEC: channels-last rows, C % 4 == 0, 16-byte aligned -- LP lanes per
// pixel move float4 groups; otherwise one thread per pixel walks the channels at the given strides.
template <bool VEC>
__global__ void __launch_bounds__(256) blend_other_bwd_kernel(const BwdParams p) {
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const int64_t total = (int64_t)HW * d.N;
  if (VEC) {
    const int C4 = d.C >> 2;
    int lp = 1;
    while (lp < 32 && lp < C4) lp <<= 1;  // lanes per pixel (power of two <= 32)
    const int lane = threadIdx.x & 31, lq = lane % lp, grp = lane / lp, G = 32 / lp;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t px0 = warp0 * G; px0 < total; px0 += nwarps * G) {
      const int64_t px = px0 + grp;
      const bool act = px < total;
      float dot = 0.f;
      float m = 1.f;
      if (act) {
        m = __ldg(p.mask + px);
        const float4* g4 = reinterpret_cast<const float4*>(p.gout) + px * C4;
        const float4* o4 = reinterpret_cast<const float4*>(p.other) + px * C4;
        float4* go4 = p.gother ? reinterpret_cast<float4*>(p.gother) + px * C4 : nullptr;
        const float om = 1.f - m;
        for (int q = lq; q < C4; q += lp) {
          const float4 g = ldg_batch(g4 + q);
          if (p.gmask) {
            const float4 o = ldg_batch(o4 + q);
            dot = fmaf(g.x, o.x, fmaf(g.y, o.y, fmaf(g.z, o.z, fmaf(g.w, o.w, dot))));
          }
          if (go4) st_stream(go4 + q, make_float4(g.x * om, g.y * om, g.z * om, g.w * om));
        }
      }
      for (int o = lp >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (act && lq == 0 && p.gmask) p.gmask[px] -= dot;
    }
  } else {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
      const int n = (int)(idx / HW);
      const int r = (int)(idx - (int64_t)n * HW);
      const int i = r / d.W, j = r - i * d.W;
      const float m = p.mask[idx];
      const int64_t gb = (int64_t)n * p.gs[0] + i * p.gs[2] + j * p.gs[3];
      float dot = 0.f;
      for (int c = 0; c < d.C; ++c) {
        const float g = p.gout[gb + c * p.gs[1]];
        if (p.gmask) dot = fmaf(g, p.other[gb + c * p.gs[1]], dot);
        if (p.gother) p.gother[gb + c * p.gs[1]] = g * (1.f - m);
      }
      if (p.gmask) p.gmask[idx] -= dot;
    }
  }
}

void launch_blend_other_bwd(const BwdParams& p, Layout lg, cudaStream_t st) {
  const Dims& d = p.d;
  const int64_t total = (int64_t)d.N * d.H * d.W;
  const bool vec = lg == LAYOUT_NHWC && (d.C & 3) == 0 && !((uintptr_t)p.gout & 15) && !((uintptr_t)p.other & 15) &&
                   !(p.gother && ((uintptr_t)p.gother & 15));
  if (vec) {
    int lp = 1;
    while (lp < 32 && lp < d.C / 4) lp <<= 1;
    const int64_t warps = (total + (32 / lp) - 1) / (32 / lp);
    int64_t blocks = (warps + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    blend_other_bwd_kernel<true><<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(p);
  } else {
    blend_other_bwd_kernel<false><<<grid_for(total), 256, 0, st>>>(p);
  }
  count_launch();
}

}  // namespace c2m
