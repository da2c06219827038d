// common.cuh -- shared device/host helpers for the C2M warp kernels (sm_100a only).
//
// Coordinate arithmetic is a bit-replica of what the reference computes on a GPU
// (SURVEY.md appendix A.3):
//   base grid     torch CPU float32 linspace(-1,1,n) (reference src/utils/ops.py:196-202)
//   normalisation flow * (1.f / (float)((n-1)/2))     (ops.py:190; ATen CUDA div-by-scalar)
//   grid add      one fp32 add                         (ops.py:191)
//   unnormalise   fma(c + 1, n, -1) * 0.5              (ATen GridSampler.cuh:22-31, align_corners=False)
//   border clip   min(n-1, max(c, 0))                  (GridSampler.cuh:55-58)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#include "../../include/c2m_warp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "c2m_b200 kernels target sm_100a only"
#endif

namespace c2m {

// Bilinear resize of the flow / mask fused into the warp.  The reference resizes both to the
// feature size right before every warp: generator.py:84-85,91-92 (F.interpolate bilinear, align_corners=False, flow
// values NOT rescaled) and utils.py:346-354 + motion_autoencoder.py:120-124 (flow: align_corners=True, values
// multiplied by new/old; mask: align_corners=False).  Arithmetic follows ATen's upsample_bilinear2d (UpSample.cuh
// area_pixel_compute_scale / area_pixel_compute_source_index, UpSampleBilinear2d.cu upsample_bilinear2d_out_frame).
struct Resize {
  int on;          // bit 0: flow is given at (fh, fw) != (H, W); bit 1: mask is given at (mh, mw) != (H, W)
  int fh, fw, mh, mw;
  int f_align;     // flow: align_corners=True and value rescale (utils.py:346-354)
  float fsy, fsx;  // flow: source-index scale (in-1)/(out-1) or in/out, float32 as ATen forms it
  float msy, msx;  // mask: in/out
  float fmulx, fmuly;  // flow value rescale: 1/(old/new) as a float32 reciprocal (ATen CUDA div-by-scalar), else 1
  int fold_t, fold_b;  // > 0: flow / mask are 5-D clips [B,c,T,h,w]; frame n = t * B + b reads plane (b, :, t)
};

// element offset of plane (frame n, channel c) of a [N,nc,h,w] tensor -- or of the 5-D clip [B,nc,T,h,w] it is a fold of
__device__ __forceinline__ int64_t plane_offset(const Resize& rs, int n, int c, int nc, int64_t hw) {
  if (rs.fold_t > 0) {
    const int t = n / rs.fold_b, b = n - t * rs.fold_b;
    return (((int64_t)b * nc + c) * rs.fold_t + t) * hw;
  }
  return ((int64_t)n * nc + c) * hw;
}

struct Dims {
  int N, C, H, W;
  int x_batch;  // distinct images in x (== N when there is no repeat)
  int padding, flags;
  float stepx, stepy;    // fp32 2/(n-1), as torch.linspace computes it
  float inv_bw, inv_bh;  // fp32 1/((n-1)/2)
  float bw, bh;          // fp32 (n-1)/2 (only for the TRUE_DIV probe variant)
  Resize rs;
};

struct FwdParams {
  Dims d;
  const float* x;
  const float* flow;
  const float* mask;
  const float* other;
  float* out;
  int64_t xs[4], os[4];
  int cchunk;  // channels per blockIdx.y slice
  int pf_tiles;  // channels-last: L2 prefetch distance in tiles (< 0: off)
};

struct BwdParams {
  Dims d;
  const float* x;
  const float* flow;
  const float* mask;
  const float* other;
  const float* gout;
  float* gx;
  float* gflow;
  float* gmask;
  float* gother;
  int64_t xs[4], gs[4];
  int cchunk;
  // deterministic mode
  long long* acc64;         // fixed-point accumulator, same element order as gx
  const unsigned* maxbits;  // bit pattern of max|gout*mask| (channels-last gather: [0] max|gout|, [1] max|mask|)
  unsigned* maxacc;         // channels-last deterministic gather: segbin_kernel forms those two maxima on its way
  int count_log2;           // ceil(log2(max contributions per destination))
  unsigned char* touched;   // deterministic channels-last gather: [x_batch*H*W] destination has terms in acc64
  int* incoh;               // deterministic channels-last gather: [x_batch] incoherent segments seen per image
  int incoh_thresh;         // more than this many: every accumulator row of the image is cleared
  // gather-form backward (contributor lists built by bin_kernel)
  int* cnt;                 // [x_batch*H*W] contributions seen per destination pixel
  void* entries;            // [x_batch*H*W][kListCap] ListEntry
  unsigned char* ovf;       // [N*H*W] bit k: corner k of this output pixel did not fit its list
  int* ovf_count;           // number of output pixels with a non-zero `ovf`
  int* ovf_list;            // their indices, in no particular order
  int n0, nframes;          // the frames [n0, n0 + nframes) a launch covers
  // channels-last local binning (no global contributor lists): candidate row segments per destination tile
  int* tcnt;                // [x_batch * tiles] segments registered per destination tile
  int2* tlist;              // [x_batch * tiles][cand_cap] segments: (index of the first pixel, live pixels <= 32)
  int cand_cap;
  int4* pixrec;             // [N*H*W] per output pixel: (x0 | y0 << 16, ax, ay, mask) -- the sampling geometry,
                            // computed once by segbin_kernel and reused by every tile that bins the pixel
  int key_mul;              // list entries name their source as pixel index * key_mul
  int pf_tiles;             // channels-last: L2 prefetch distance in tiles (< 0: off)
  int64_t ovf_stride;       // channel-sliced gather: slice s flags its own list overflows in ovf[(s + 1) * ovf_stride + pixel]
                            // and lists them as pixel | (s + 1) << 24 (tag 0: all channels, from segbin_kernel)
  float* gpart;             // channel-sliced gather (small levels): [slices][gflow N*2*HW | gmask N*HW] partial sums
  // incoherent flows (a row segment whose samples spread over more than 12 destination tiles): its pixels are
  // registered one by one with the destination tiles they touch (counting sort keyed by tile: count in segbin_kernel,
  // flex_scan_kernel, flex_fill_kernel) and those tiles' grad-input is formed by gather_flex_kernel
  int* bcount;              // [x_batch * tiles] pixel registrations per destination tile (0: the tile is not "flex")
  int* bstart;              // [x_batch * tiles] start of the tile's registrations in `pool`
  int* bfill;               // [x_batch * tiles] fill cursor
  int* pool;                // [4 * N*H*W] registered output pixels, grouped by destination tile
  int* iseg_count;          // number of incoherent segments ...
  int* iseg_list;           // ... and their ids (frame * segments per frame + segment)
  int* flex_count;          // number of flex tiles, their ids, and the work-queue cursor of gather_flex_kernel
  int* flex_list;
  int* flex_next;
};

// One output sample of ATen's upsample_bilinear2d: plane `pl` [Hs, Ws] sampled for output pixel (i, j).
__device__ __forceinline__ void resize_taps(float scale, int dst, int in_size, bool align, int& i0, int& ip, float& l0,
                                            float& l1) {
  float r;
  if (align) {
    r = scale * (float)dst;
  } else {
    r = scale * ((float)dst + 0.5f) - 0.5f;
    r = r < 0.f ? 0.f : r;
  }
  i0 = (int)r;
  ip = (i0 < in_size - 1) ? 1 : 0;
  l1 = r - (float)i0;
  l0 = 1.f - l1;
}
__device__ __forceinline__ float bilerp_plane(const float* __restrict__ pl, int Hs, int Ws, float sy, float sx, bool align,
                                              int i, int j) {
  int h1, h1p, w1, w1p;
  float h0l, h1l, w0l, w1l;
  resize_taps(sy, i, Hs, align, h1, h1p, h0l, h1l);
  resize_taps(sx, j, Ws, align, w1, w1p, w0l, w1l);
  const float* r0 = pl + (int64_t)h1 * Ws + w1;
  const float* r1 = r0 + (int64_t)h1p * Ws;
  const float a = __ldg(r0), b = __ldg(r0 + w1p), c = __ldg(r1), e = __ldg(r1 + w1p);
  // the expression of upsample_bilinear2d_out_frame, left to the compiler's default contraction as in ATen's build
  return h0l * (w0l * a + w1l * b) + h1l * (w0l * c + w1l * e);
}

// flow (fx, fy) and mask value of output pixel (n, i, j): read directly, or resized on the fly from the tensors the
// caller holds at another resolution.  `flow` / `mask` are the pointers as passed to the entry point.
__device__ __forceinline__ void fetch_flow_mask(const Dims& d, const float* __restrict__ flow,
                                                const float* __restrict__ mask, int n, int i, int j, float& fx, float& fy,
                                                float& m) {
  const Resize& rs = d.rs;
  if (rs.on & 1) {
    const int64_t fhw = (int64_t)rs.fh * rs.fw;
    const float* f0 = flow + plane_offset(rs, n, 0, 2, fhw);
    const float* f1 = flow + plane_offset(rs, n, 1, 2, fhw);
    if (rs.fh == d.H && rs.fw == d.W) {  // folded clip at the feature size: a plain read (ATen copies in that case)
      fx = __ldg(f0 + i * d.W + j);
      fy = __ldg(f1 + i * d.W + j);
    } else {
      fx = bilerp_plane(f0, rs.fh, rs.fw, rs.fsy, rs.fsx, rs.f_align != 0, i, j);
      fy = bilerp_plane(f1, rs.fh, rs.fw, rs.fsy, rs.fsx, rs.f_align != 0, i, j);
    }
    if (rs.f_align) {
      fx = __fmul_rn(fx, rs.fmulx);
      fy = __fmul_rn(fy, rs.fmuly);
    }
  } else {
    const float* fl = flow + (int64_t)n * 2 * d.H * d.W + i * d.W + j;
    fx = __ldg(fl);
    fy = __ldg(fl + d.H * d.W);
  }
  if (mask) {
    if (rs.on & 2) {
      const float* m0 = mask + plane_offset(rs, n, 0, 1, (int64_t)rs.mh * rs.mw);
      m = (rs.mh == d.H && rs.mw == d.W) ? __ldg(m0 + i * d.W + j)
                                         : bilerp_plane(m0, rs.mh, rs.mw, rs.msy, rs.msx, false, i, j);
    } else {
      m = __ldg(mask + (int64_t)n * d.H * d.W + i * d.W + j);
    }
  } else {
    m = 1.f;
  }
}

// Per-pixel sampling geometry, shared by forward and backward.
struct Geo {
  float wnw, wne, wsw, wse;  // bilinear weights (ATen order nw, ne, sw, se)
  int x0, y0, x1, y1;        // corner indices, clamped into the image (loads are always legal)
  bool oknw, okne, oksw, okse;
  float gmx, gmy;            // d(clipped coord)/d(flow) = size/2 * clip-grad * inv_b
  float ax, ay;              // ix - x0, iy - y0 (for the coordinate gradient)
  bool clipx, clipy;         // border padding: coordinate was clipped (its gradient w.r.t. the flow is zero)
};

__device__ __forceinline__ float base_coord(int k, int n, float step) {
  // torch CPU linspace: first half fma(step,k,-1), second half fma(-step,n-1-k,1); n==1 -> -1
  if (n == 1) return -1.f;
  return (k < (n >> 1)) ? fmaf(step, (float)k, -1.f) : fmaf(-step, (float)(n - 1 - k), 1.f);
}

// PROBE keeps the run-time switches of tools/probe.py (true division / unfused multiply-add); the layout-
// specialised kernels compile them out (launchers send such calls to the generic kernels).
template <bool PROBE>
__device__ __forceinline__ float unnormalized(float f, int k, int n, float step, float inv_b, float b,
                                              int flags) {
  const float g = base_coord(k, n, step);
  // intrinsics keep nvcc from contracting across the reference's materialised temporaries
  const float nf = (PROBE && (flags & C2M_FLAG_TRUE_DIV)) ? __fdiv_rn(f, b) : __fmul_rn(f, inv_b);
  const float c1 = __fadd_rn(__fadd_rn(g, nf), 1.f);
  // align_corners=True (ATen grid_sampler_unnormalize): ((coord + 1) / 2) * (size - 1)
  if (PROBE && (flags & C2M_FLAG_ALIGN_CORNERS)) return __fmul_rn(__fmul_rn(c1, 0.5f), (float)(n - 1));
  const float u = (PROBE && (flags & C2M_FLAG_NO_FMA)) ? __fadd_rn(__fmul_rn(c1, (float)n), -1.f)
                                                       : fmaf(c1, (float)n, -1.f);
  return u * 0.5f;
}

// BWD selects ATen's clip_coordinates_set_grad behaviour (NaN is not clipped there and ends up at
// -100 through safe_downgrade_to_int_range, GridSampler.cuh:64-82,141-148).
template <bool BWD, bool PROBE = false>
__device__ __forceinline__ void make_geo(const Dims& d, float fx, float fy, int i, int j, Geo& g) {
  float ix, iy;
  float sx = d.inv_bw, sy = d.inv_bh;
  if (PROBE && (d.flags & C2M_FLAG_COORD_GRID)) {
    // (fx, fy) is a normalised sampling location (reference utils.grid_sample, ops.py:183-184)
    if (d.flags & C2M_FLAG_ALIGN_CORNERS) {
      ix = __fmul_rn(__fmul_rn(__fadd_rn(fx, 1.f), 0.5f), (float)(d.W - 1));
      iy = __fmul_rn(__fmul_rn(__fadd_rn(fy, 1.f), 0.5f), (float)(d.H - 1));
    } else {
      ix = fmaf(__fadd_rn(fx, 1.f), (float)d.W, -1.f) * 0.5f;
      iy = fmaf(__fadd_rn(fy, 1.f), (float)d.H, -1.f) * 0.5f;
    }
    sx = 1.f;
    sy = 1.f;
  } else {
    ix = unnormalized<PROBE>(fx, j, d.W, d.stepx, d.inv_bw, d.bw, d.flags);
    iy = unnormalized<PROBE>(fy, i, d.H, d.stepy, d.inv_bh, d.bh, d.flags);
  }
  float cgx = 1.f, cgy = 1.f;
  if (d.padding == C2M_PAD_BORDER) {
    if (BWD) {
      cgx = (ix > 0.f && ix < (float)(d.W - 1)) ? 1.f : 0.f;
      cgy = (iy > 0.f && iy < (float)(d.H - 1)) ? 1.f : 0.f;
      ix = (ix != ix) ? -100.f : fminf((float)(d.W - 1), fmaxf(ix, 0.f));
      iy = (iy != iy) ? -100.f : fminf((float)(d.H - 1), fmaxf(iy, 0.f));
    } else {
      ix = fminf((float)(d.W - 1), fmaxf(ix, 0.f));  // fmaxf(NaN,0)=0, as ATen's ::max
      iy = fminf((float)(d.H - 1), fmaxf(iy, 0.f));
    }
  } else {
    if (!(ix <= 2147483646.f && ix >= -2147483648.f)) ix = -100.f;  // also catches NaN / inf
    if (!(iy <= 2147483646.f && iy >= -2147483648.f)) iy = -100.f;
  }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
  g.wnw = (fx1 - ix) * (fy1 - iy);
  g.wne = (ix - fx0) * (fy1 - iy);
  g.wsw = (fx1 - ix) * (iy - fy0);
  g.wse = (ix - fx0) * (iy - fy0);
  g.ax = ix - fx0;
  g.ay = iy - fy0;
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const bool x0ok = (x0 >= 0) & (x0 < d.W), x1ok = (x1 >= 0) & (x1 < d.W);
  const bool y0ok = (y0 >= 0) & (y0 < d.H), y1ok = (y1 >= 0) & (y1 < d.H);
  g.oknw = x0ok & y0ok;
  g.okne = x1ok & y0ok;
  g.oksw = x0ok & y1ok;
  g.okse = x1ok & y1ok;
  g.x0 = min(max(x0, 0), d.W - 1);
  g.x1 = min(max(x1, 0), d.W - 1);
  g.y0 = min(max(y0, 0), d.H - 1);
  g.y1 = min(max(y1, 0), d.H - 1);
  // ATen: grad_grid = (size/2 * clip_grad) * gix; autograd of ops.py:190 multiplies by 1/((size-1)/2)
  // (align_corners=True: (size - 1) / 2, grid_sampler_unnormalize_set_grad)
  const bool align = PROBE && (d.flags & C2M_FLAG_ALIGN_CORNERS);
  g.gmx = cgx * (0.5f * (float)(align ? d.W - 1 : d.W)) * sx;
  g.gmy = cgy * (0.5f * (float)(align ? d.H - 1 : d.H)) * sy;
  g.clipx = cgx == 0.f;
  g.clipy = cgy == 0.f;
}

// ---------------------------------------------------------------------------------------------
// Deterministic grad-input: every term w * gout is converted to 64-bit fixed point with a power-of-two
// scale and summed as an integer -- the result does not depend on the order (or on which mechanism adds
// which term) and is converted back to float once.  scale = 2^(60 - count_log2 - exponent(max |term|)).
__device__ __forceinline__ float fixed_scale_from(float mx, int count_log2) {
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &e);
  int k = 60 - count_log2 - e;
  k = max(-120, min(120, k));
  return exp2f((float)k);
}
__device__ __forceinline__ long long to_fixed(float term, float scale) { return __float2ll_rn(term * scale); }
// 1 / scale for the one conversion back to float -- or NaN when the upstream gradient holds a non-finite value (the
// max-reductions report it as +inf): an integer sum cannot carry NaN / inf, so the destinations summed that way are
// poisoned as a whole instead of coming back finite (ATen and the float paths propagate non-finite terms).
__device__ __forceinline__ float fixed_inv_scale(float mx, float scale) {
  return (mx < 3.0e38f) ? 1.f / scale : __int_as_float(0x7fc00000);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "C2M_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra C2M_DONE;\n"
      "bra C2M_WAIT;\n"
      "C2M_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 3-D tiled tensor load: coordinates are (innermost, middle, outermost) element indices.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// Bulk L2 prefetch (TMA engine, no destination): `bytes` is a multiple of 16, `p` 16-byte aligned.
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// streaming (evict-first) accesses for data touched exactly once
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// Read-only 128-bit load as a volatile asm statement: the compiler keeps a run of these together
// (it otherwise interleaves loads with their first uses and leaves only two in flight), which is
// what gives the gather kernels their memory-level parallelism.
__device__ __forceinline__ float4 ldg_batch(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Predicated flavour: the load is issued only when `on` is set, otherwise the result is zero (no
// branch, no memory transaction).
__device__ __forceinline__ float4 ldg_batch_if(const float4* p, bool on) {
  float4 v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %5, 0;\n"
      "mov.f32 %0, 0f00000000;\n"
      "mov.f32 %1, 0f00000000;\n"
      "mov.f32 %2, 0f00000000;\n"
      "mov.f32 %3, 0f00000000;\n"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n"
      "}\n"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"((int)on));
  return v;
}
__device__ __forceinline__ float ldg_batch(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Tile walk shared by the layout-specialised kernels: persistent CTAs step through (n, tile_y,
// tile_x) work items; the tile's two flow planes and its mask plane are staged in shared memory
// by TMA one tile ahead (double buffered, one mbarrier per buffer).
// ---------------------------------------------------------------------------------------------
template <int TH, int TW>
struct TileSmem {
  alignas(128) float flow[2][2][TH][TW];
  alignas(128) float mask[2][TH][TW];
  alignas(8) uint64_t bar[2];
};

template <int TH, int TW, bool HAS_MASK>
__device__ __forceinline__ void issue_tile(TileSmem<TH, TW>& s, const CUtensorMap* tmf, const CUtensorMap* tmm,
                                           int t, int tiles_x, int tiles_y, int b) {
  const int bx = t % tiles_x;
  const int r = t / tiles_x;
  const int by = r % tiles_y;
  const int n = r / tiles_y;
  constexpr uint32_t bytes = (HAS_MASK ? 3u : 2u) * TH * TW * sizeof(float);
  mbar_expect_tx(&s.bar[b], bytes);
  tma_load_3d(&s.flow[b][0][0][0], tmf, &s.bar[b], bx * TW, by * TH, n * 2);
  if (HAS_MASK) tma_load_3d(&s.mask[b][0][0], tmm, &s.bar[b], bx * TW, by * TH, n);
}

template <int TH, int TW, bool HAS_MASK>
__device__ __forceinline__ void tile_pipeline_init(TileSmem<TH, TW>& s, const CUtensorMap* tmf,
                                                   const CUtensorMap* tmm, int first_tile, int total, int tiles_x,
                                                   int tiles_y) {
  if (threadIdx.x == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(tmf);
    if (HAS_MASK) tma_prefetch_desc(tmm);
  }
  __syncthreads();
  if (threadIdx.x == 0 && first_tile < total) issue_tile<TH, TW, HAS_MASK>(s, tmf, tmm, first_tile, tiles_x, tiles_y, 0);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
enum Layout { LAYOUT_NCHW = 0, LAYOUT_NHWC = 1, LAYOUT_OTHER = 2 };

// Gather-form backward: contributor lists (one per destination pixel of grad-input)
constexpr int kSplitTiles = 1184;  // fewer tiles than this (8 per SM): slice the channels over blockIdx.y
constexpr int kSplitMax = 8;
constexpr int kListCap = 8;   // global lists (NCHW path): in-line entries per destination; the tail goes through atomics
constexpr int kLocalCap = 12; // channels-last local binning: entries per destination in shared memory
constexpr int kLocalCapDet = 14;  // ... in deterministic mode (14: the CTA then uses exactly 40 960 B, see gather_nhwc_kernel)
constexpr int kCandPerFrame = 96;  // candidate row segments a destination tile can register per source frame
constexpr int kCandMax = 256;      // ... and in total (8 warps x 32 lanes preload the ids)
struct ListEntry {
  int src;    // source (output) pixel n*H*W + i*W + j, times BwdParams::key_mul
  float w;    // bilinear weight * mask
};

void set_error(const char* fmt, ...);
// measurement hook (c2m_warp_profile): record the thread's event pair around the dominant kernel of a call
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);
void count_launch(int n = 1);
int sm_count();
// L2 prefetch distance of the channels-last kernels, in tiles (< 0: off); the environment variable
// C2M_WARP_PREFETCH_TILES overrides the kernel's default (tuning hook)
// sizes + the fp32 coordinate constants of the reference's grid arithmetic (validates; sets the error text)
int fill_dims(Dims& d, int64_t N, int C, int H, int W, int64_t x_batch, int padding, int flags);
int prefetch_tiles(int dflt);
// number of blockIdx.y channel slices the channels-last kernels use for a level (1: none)
int channel_slices(int64_t N, int C, int H, int W);
// Number of CTAs of `kernel` (static shared memory only) that are resident on the whole device at once:
// the grid size of a persistent launch.
int resident_ctas(const void* kernel, int threads);
// Encodes a [planes, H, W] float32 tensor map with box (box_w, box_h, box_p); returns false when
// the tensor does not satisfy TMA's alignment rules (caller falls back to plain loads).
bool make_tensor_map_3d(CUtensorMap* tm, const float* base, int W, int H, int64_t planes, int box_w, int box_h,
                        int box_p);

struct TileMaps {
  CUtensorMap flow, mask;
  bool ok;  // false: alignment rules not met (or C2M_FLAG_NO_TMA) -> kernels use plain loads
};
TileMaps make_tile_maps(const Dims& d, const float* flow, const float* mask, int TH, int TW);

int launch_fwd(const FwdParams& p, Layout lx, Layout lo, cudaStream_t st);
// fused-resize backward helpers (warp_resize.cu): materialise the resized flow / mask; back-propagate their gradients
int fill_resize(Dims& d, const c2m_resize* rs);
void launch_resize_fwd(const Dims& d, const float* flow_src, const float* mask_src, float* flow_out, float* mask_out,
                       cudaStream_t st);
void launch_resize_bwd(const Dims& d, const float* gflow_small, const float* gmask_small, float* gflow_src,
                       float* gmask_src, cudaStream_t st);
int launch_bwd(const BwdParams& p, Layout lx, Layout lg, void* workspace, size_t workspace_bytes, cudaStream_t st);
void launch_relayout(const float* in, float* out, int64_t n, int C, int HW, bool to_nhwc, cudaStream_t st);
size_t plan_bytes(const Dims& d);
int launch_plan(const BwdParams& p, void* plan, size_t bytes, cudaStream_t st);
size_t bwd_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx, int flags);

}  // namespace c2m
