// warp_fwd.cu -- fused forward: flow -> sampling coordinates, bilinear gather, occlusion blend.
// Replaces reference src/utils/ops.py:187-202 + src/modules/generator/generator.py:93 (one kernel
// instead of CPU grid build + H2D + div/div/cat/add + grid_sampler_2d + mul).
//
// Kernels
//   fwd_generic_kernel  any strides / any option; one thread per pixel, scalar channel loop.
//   fwd_nchw_kernel     NCHW-contiguous; persistent CTAs walk (n, tile) work items, the tile's two
//                       flow planes and the mask plane are staged in shared memory by TMA
//                       (cp.async.bulk.tensor.3d, double buffered, mbarrier-signalled) one tile
//                       ahead; one thread per pixel, channel loop unrolled for memory-level
//                       parallelism; stores are 128-byte coalesced along W.
//   fwd_nhwc_kernel     channels-last; same tile walk and TMA staging; LP lanes per pixel, each
//                       moving float4 channel groups: four 128-bit corner loads + one 128-bit store.
#include "common.cuh"

namespace c2m {

// ---------------------------------------------------------------------------------------------
template <bool HAS_OTHER>
__global__ void __launch_bounds__(256) fwd_generic_kernel(const FwdParams p) {
  const Dims& d = p.d;
  const int64_t HW = (int64_t)d.H * d.W;
  const int64_t total = HW * d.N;
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
    make_geo<false>(d, fx, fy, i, j, g);
    const float* xb = p.x + (int64_t)(n % d.x_batch) * p.xs[0];
    const int64_t onw = g.y0 * p.xs[2] + g.x0 * p.xs[3], one = g.y0 * p.xs[2] + g.x1 * p.xs[3];
    const int64_t osw = g.y1 * p.xs[2] + g.x0 * p.xs[3], ose = g.y1 * p.xs[2] + g.x1 * p.xs[3];
    const int64_t ob = (int64_t)n * p.os[0] + i * p.os[2] + j * p.os[3];
    for (int c = 0; c < d.C; ++c) {
      const float* xc = xb + c * p.xs[1];
      const float vnw = g.oknw ? xc[onw] : 0.f, vne = g.okne ? xc[one] : 0.f;
      const float vsw = g.oksw ? xc[osw] : 0.f, vse = g.okse ? xc[ose] : 0.f;
      float acc = vnw * g.wnw;
      acc = fmaf(vne, g.wne, acc);
      acc = fmaf(vsw, g.wsw, acc);
      acc = fmaf(vse, g.wse, acc);
      float o = p.mask ? __fmul_rn(acc, m) : acc;
      if (HAS_OTHER) o = __fadd_rn(o, __fmul_rn(1.f - m, p.other[ob + c * p.os[1]]));
      p.out[ob + c * p.os[1]] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tile walker shared by the NCHW and NHWC kernels: stages flow (2 planes) and mask (1 plane) of
// tile t into buffer b.
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

// ---------------------------------------------------------------------------------------------
template <int TH, int TW, bool HAS_MASK, bool USE_TMA, int UNROLL>
__global__ void __launch_bounds__(TH* TW, (TH * TW <= 256) ? 768 / (TH * TW) : 2) fwd_nchw_kernel(const __grid_constant__ FwdParams p,
                                                          const __grid_constant__ CUtensorMap tm_flow,
                                                          const __grid_constant__ CUtensorMap tm_mask) {
  __shared__ TileSmem<TH, TW> s;
  const Dims& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid % TW, ty = tid / TW;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int total = d.N * tiles_y * tiles_x;
  const int c0 = blockIdx.y * p.cchunk;
  const int nc = min(p.cchunk, d.C - c0);
  const int HW = d.H * d.W;

  if (USE_TMA) {
    if (tid == 0) {
      mbar_init(&s.bar[0], 1);
      mbar_init(&s.bar[1], 1);
      mbar_fence_init();
      tma_prefetch_desc(&tm_flow);
      if (HAS_MASK) tma_prefetch_desc(&tm_mask);
    }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total)
      issue_tile<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, blockIdx.x, tiles_x, tiles_y, 0);
  }
  int buf = 0;
  uint32_t phases = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int bx = t % tiles_x;
    const int r = t / tiles_x;
    const int by = r % tiles_y;
    const int n = r / tiles_y;
    const int i = by * TH + ty, j = bx * TW + tx;
    const bool live = (i < d.H) & (j < d.W);
    float fx = 0.f, fy = 0.f, m = 1.f;
    if (USE_TMA) {
      const int tn = t + gridDim.x;
      if (tid == 0 && tn < total) issue_tile<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, tn, tiles_x, tiles_y, buf ^ 1);
      mbar_wait(&s.bar[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      fx = s.flow[buf][0][ty][tx];
      fy = s.flow[buf][1][ty][tx];
      if (HAS_MASK) m = s.mask[buf][ty][tx];
    } else if (live) {
      const float* fl = p.flow + (int64_t)n * 2 * HW + i * d.W + j;
      fx = __ldg(fl);
      fy = __ldg(fl + HW);
      if (HAS_MASK) m = __ldg(p.mask + (int64_t)n * HW + i * d.W + j);
    }
    if (live) {
      Geo g;
      make_geo<false>(d, fx, fy, i, j, g);
      const int onw = g.y0 * d.W + g.x0, one = g.y0 * d.W + g.x1;
      const int osw = g.y1 * d.W + g.x0, ose = g.y1 * d.W + g.x1;
      const float* xc = p.x + ((int64_t)(n % d.x_batch) * d.C + c0) * HW;
      const float* pnw = xc + onw;
      const float* pne = xc + one;
      const float* psw = xc + osw;
      const float* pse = xc + ose;
      float* oc = p.out + ((int64_t)n * d.C + c0) * HW + i * d.W + j;
#pragma unroll UNROLL
      for (int c = 0; c < nc; ++c) {
        float vnw = __ldg(pnw), vne = __ldg(pne);
        float vsw = __ldg(psw), vse = __ldg(pse);
        vnw = g.oknw ? vnw : 0.f;
        vne = g.okne ? vne : 0.f;
        vsw = g.oksw ? vsw : 0.f;
        vse = g.okse ? vse : 0.f;
        float acc = vnw * g.wnw;
        acc = fmaf(vne, g.wne, acc);
        acc = fmaf(vsw, g.wsw, acc);
        acc = fmaf(vse, g.wse, acc);
        st_stream(oc, HAS_MASK ? __fmul_rn(acc, m) : acc);
        pnw += HW;
        pne += HW;
        psw += HW;
        pse += HW;
        oc += HW;
      }
    }
    if (USE_TMA) {
      __syncthreads();  // every thread has consumed buffer `buf` before it is refilled
      buf ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// channels-last: LP lanes cooperate on one pixel, lane q handles float4 groups q, q+LP, ...
template <int TH, int TW, int LP, bool HAS_MASK, bool USE_TMA>
__global__ void __launch_bounds__(256) fwd_nhwc_kernel(const __grid_constant__ FwdParams p,
                                                       const __grid_constant__ CUtensorMap tm_flow,
                                                       const __grid_constant__ CUtensorMap tm_mask) {
  __shared__ TileSmem<TH, TW> s;
  constexpr int NT = 256;
  constexpr int PIX_PER_PASS = NT / LP;
  const Dims& d = p.d;
  const int tid = threadIdx.x;
  const int lane_q = tid % LP, pix0 = tid / LP;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int total = d.N * tiles_y * tiles_x;
  const int HW = d.H * d.W;
  const int C4 = d.C >> 2;
  const int q0 = blockIdx.y * p.cchunk;  // cchunk counted in float4 groups here
  const int q1 = min(C4, q0 + p.cchunk);

  if (USE_TMA) {
    if (tid == 0) {
      mbar_init(&s.bar[0], 1);
      mbar_init(&s.bar[1], 1);
      mbar_fence_init();
      tma_prefetch_desc(&tm_flow);
      if (HAS_MASK) tma_prefetch_desc(&tm_mask);
    }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < total)
      issue_tile<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, blockIdx.x, tiles_x, tiles_y, 0);
  }
  int buf = 0;
  uint32_t phases = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int bx = t % tiles_x;
    const int r = t / tiles_x;
    const int by = r % tiles_y;
    const int n = r / tiles_y;
    if (USE_TMA) {
      const int tn = t + gridDim.x;
      if (tid == 0 && tn < total) issue_tile<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, tn, tiles_x, tiles_y, buf ^ 1);
      mbar_wait(&s.bar[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
    }
    const float4* xb = reinterpret_cast<const float4*>(p.x + (int64_t)(n % d.x_batch) * HW * d.C);
    float4* ob = reinterpret_cast<float4*>(p.out + (int64_t)n * HW * d.C);
#pragma unroll 2
    for (int pp = pix0; pp < TH * TW; pp += PIX_PER_PASS) {
      const int ty = pp / TW, tx = pp % TW;
      const int i = by * TH + ty, j = bx * TW + tx;
      if ((i >= d.H) | (j >= d.W)) continue;
      float fx, fy, m = 1.f;
      if (USE_TMA) {
        fx = s.flow[buf][0][ty][tx];
        fy = s.flow[buf][1][ty][tx];
        if (HAS_MASK) m = s.mask[buf][ty][tx];
      } else {
        const float* fl = p.flow + (int64_t)n * 2 * HW + i * d.W + j;
        fx = __ldg(fl);
        fy = __ldg(fl + HW);
        if (HAS_MASK) m = __ldg(p.mask + (int64_t)n * HW + i * d.W + j);
      }
      Geo g;
      make_geo<false>(d, fx, fy, i, j, g);
      const float4* pnw = xb + (int64_t)(g.y0 * d.W + g.x0) * C4;
      const float4* pne = xb + (int64_t)(g.y0 * d.W + g.x1) * C4;
      const float4* psw = xb + (int64_t)(g.y1 * d.W + g.x0) * C4;
      const float4* pse = xb + (int64_t)(g.y1 * d.W + g.x1) * C4;
      float4* po = ob + (int64_t)(i * d.W + j) * C4;
      const float wnw = g.oknw ? g.wnw : 0.f, wne = g.okne ? g.wne : 0.f;
      const float wsw = g.oksw ? g.wsw : 0.f, wse = g.okse ? g.wse : 0.f;
#pragma unroll 2
      for (int q = q0 + lane_q; q < q1; q += LP) {
        float4 a = __ldg(pnw + q), b = __ldg(pne + q), c = __ldg(psw + q), e = __ldg(pse + q);
        if (!g.oknw) a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!g.okne) b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!g.oksw) c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!g.okse) e = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = fmaf(e.x, wse, fmaf(c.x, wsw, fmaf(b.x, wne, a.x * wnw)));
        o.y = fmaf(e.y, wse, fmaf(c.y, wsw, fmaf(b.y, wne, a.y * wnw)));
        o.z = fmaf(e.z, wse, fmaf(c.z, wsw, fmaf(b.z, wne, a.z * wnw)));
        o.w = fmaf(e.w, wse, fmaf(c.w, wsw, fmaf(b.w, wne, a.w * wnw)));
        if (HAS_MASK) {
          o.x = __fmul_rn(o.x, m);
          o.y = __fmul_rn(o.y, m);
          o.z = __fmul_rn(o.z, m);
          o.w = __fmul_rn(o.w, m);
        }
        st_stream(po + q, o);
      }
    }
    if (USE_TMA) {
      __syncthreads();
      buf ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static int pick_grid_x(int total_tiles, int ctas_per_sm, int ysplit) {
  const int cap = sm_count() * ctas_per_sm;
  int gx = cap / ysplit;
  if (gx < 1) gx = 1;
  return total_tiles < gx ? total_tiles : gx;
}

template <int TH, int TW, int UNROLL>
static int launch_nchw_t(FwdParams p, cudaStream_t st) {
  const Dims& d = p.d;
  const int tiles = d.N * ((d.H + TH - 1) / TH) * ((d.W + TW - 1) / TW);
  // split channels across blockIdx.y only when there are too few tiles to fill the machine
  int ysplit = 1;
  const int want = sm_count() * 4;
  while (tiles * ysplit < want && (d.C / (ysplit * 2)) >= 8) ysplit *= 2;
  p.cchunk = (d.C + ysplit - 1) / ysplit;
  ysplit = (d.C + p.cchunk - 1) / p.cchunk;
  CUtensorMap tmf, tmm;
  memset(&tmf, 0, sizeof(tmf));
  memset(&tmm, 0, sizeof(tmm));
  bool tma = !(d.flags & C2M_FLAG_NO_TMA) && make_tensor_map_3d(&tmf, p.flow, d.W, d.H, (int64_t)d.N * 2, TW, TH, 2);
  if (tma && p.mask) tma = make_tensor_map_3d(&tmm, p.mask, d.W, d.H, d.N, TW, TH, 1);
  constexpr int NT = TH * TW;
  const int per_sm = 2048 / NT;
  dim3 grid(pick_grid_x(tiles, per_sm, ysplit), ysplit);
#define C2M_LAUNCH(MASK, TMA) \
  fwd_nchw_kernel<TH, TW, MASK, TMA, UNROLL><<<grid, NT, 0, st>>>(p, tmf, tmm)
  if (p.mask) {
    if (tma) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tma) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
  return C2M_OK;
}

template <int LP>
static int launch_nhwc_t(FwdParams p, cudaStream_t st) {
  constexpr int TH = 4, TW = 32;
  const Dims& d = p.d;
  const int tiles = d.N * ((d.H + TH - 1) / TH) * ((d.W + TW - 1) / TW);
  const int C4 = d.C / 4;
  int ysplit = 1;
  const int want = sm_count() * 4;
  while (tiles * ysplit < want && (C4 / (ysplit * 2)) >= LP) ysplit *= 2;
  p.cchunk = (C4 + ysplit - 1) / ysplit;
  ysplit = (C4 + p.cchunk - 1) / p.cchunk;
  CUtensorMap tmf, tmm;
  memset(&tmf, 0, sizeof(tmf));
  memset(&tmm, 0, sizeof(tmm));
  bool tma = !(d.flags & C2M_FLAG_NO_TMA) && make_tensor_map_3d(&tmf, p.flow, d.W, d.H, (int64_t)d.N * 2, TW, TH, 2);
  if (tma && p.mask) tma = make_tensor_map_3d(&tmm, p.mask, d.W, d.H, d.N, TW, TH, 1);
  dim3 grid(pick_grid_x(tiles, 8, ysplit), ysplit);
#define C2M_LAUNCH(MASK, TMA) fwd_nhwc_kernel<TH, TW, LP, MASK, TMA><<<grid, 256, 0, st>>>(p, tmf, tmm)
  if (p.mask) {
    if (tma) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tma) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
  return C2M_OK;
}

int launch_fwd(const FwdParams& p, Layout lx, Layout lo, cudaStream_t st) {
  const Dims& d = p.d;
  const bool generic = (d.flags & (C2M_FLAG_FORCE_GENERIC | C2M_FLAG_COORD_GRID)) || p.other != nullptr || lx != lo || lx == LAYOUT_OTHER ||
                       (int64_t)d.H * d.W >= (1ll << 30);
  if (!generic && lx == LAYOUT_NCHW) {
    const int variant = (d.flags >> 16) & 0xf;  // tuning hook (bench sweeps); 0 = default
    switch (variant) {
      case 1: return launch_nchw_t<4, 64, 8>(p, st);
      case 2: return launch_nchw_t<8, 64, 8>(p, st);
      case 3: return launch_nchw_t<16, 32, 8>(p, st);
      case 4: return launch_nchw_t<8, 32, 4>(p, st);
      case 5: return launch_nchw_t<4, 32, 8>(p, st);
      default: return launch_nchw_t<8, 32, 8>(p, st);
    }
  }
  if (!generic && lx == LAYOUT_NHWC && (d.C % 4) == 0 && ((uintptr_t)p.x % 16) == 0 && ((uintptr_t)p.out % 16) == 0) {
    const int C4 = d.C / 4;
    if (C4 >= 16) return launch_nhwc_t<16>(p, st);
    if (C4 >= 8) return launch_nhwc_t<8>(p, st);
    if (C4 >= 4) return launch_nhwc_t<4>(p, st);
    if (C4 >= 2) return launch_nhwc_t<2>(p, st);
    return launch_nhwc_t<1>(p, st);
  }
  const int64_t total = (int64_t)d.N * d.H * d.W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (p.other)
    fwd_generic_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(p);
  else
    fwd_generic_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(p);
  count_launch();
  return C2M_OK;
}

}  // namespace c2m
