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
//   fwd_nhwc_kernel     channels-last; one CTA per 8 x 32 tile (non-persistent), flow/mask tile by TMA;
//                       warp w owns tile row w: lane t writes the geometry of pixel t to the warp's
//                       shared-memory slice, then LP lanes per pixel move float4 channel groups: four
//                       128-bit corner loads + one 128-bit store, every access a full 16-byte-per-lane
//                       coalesced segment whatever the flow does.
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
    float fx, fy, m;
    if (gridmode) {
      const float* fl = p.flow + ((int64_t)n * HW + r) * 2;
      fx = fl[0];
      fy = fl[1];
      m = p.mask ? p.mask[(int64_t)n * HW + r] : 1.f;
    } else {
      fetch_flow_mask(d, p.flow, p.mask, n, i, j, fx, fy, m);
    }
    Geo g;
    make_geo<false, true>(d, fx, fy, i, j, g);
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
template <int TH, int TW, bool HAS_MASK, bool USE_TMA, int UNROLL>
__global__ void __launch_bounds__(TH* TW, (TH * TW <= 256) ? 768 / (TH * TW) : 2)
    fwd_nchw_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ CUtensorMap tm_flow,
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

  if (USE_TMA) tile_pipeline_init<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, blockIdx.x, total, tiles_x, tiles_y);
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
      fetch_flow_mask(d, p.flow, HAS_MASK ? p.mask : nullptr, n, i, j, fx, fy, m);
    }
    if (live) {
      Geo g;
      make_geo<false>(d, fx, fy, i, j, g);
      const float* xc = p.x + ((int64_t)(n % d.x_batch) * d.C + c0) * HW;
      const float* pnw = xc + (g.y0 * d.W + g.x0);
      const float* pne = xc + (g.y0 * d.W + g.x1);
      const float* psw = xc + (g.y1 * d.W + g.x0);
      const float* pse = xc + (g.y1 * d.W + g.x1);
      float* oc = p.out + ((int64_t)n * d.C + c0) * HW + i * d.W + j;
      // blend operand (out = m*warp + (1-m)*other): read at the output position, same strides as `out`.  Its own
      // copy of the channel loop: the plain loop stays exactly as it was (it is latency-bound on its four gathers
      // per channel and loses a third of its speed to any extra work in the body)
      const float* otc = (HAS_MASK && p.other) ? p.other + ((int64_t)n * d.C + c0) * HW + i * d.W + j : nullptr;
      if (otc) {
        const float om = 1.f - m;
#pragma unroll UNROLL
        for (int c = 0; c < nc; ++c) {
          float vnw = __ldg(pnw), vne = __ldg(pne);
          float vsw = __ldg(psw), vse = __ldg(pse);
          const float ot = __ldg(otc);
          vnw = g.oknw ? vnw : 0.f;
          vne = g.okne ? vne : 0.f;
          vsw = g.oksw ? vsw : 0.f;
          vse = g.okse ? vse : 0.f;
          float acc = vnw * g.wnw;
          acc = fmaf(vne, g.wne, acc);
          acc = fmaf(vsw, g.wsw, acc);
          acc = fmaf(vse, g.wse, acc);
          st_stream(oc, __fadd_rn(__fmul_rn(acc, m), __fmul_rn(om, ot)));
          pnw += HW;
          pne += HW;
          psw += HW;
          pse += HW;
          oc += HW;
          otc += HW;
        }
      } else if (nc == 3) {
        // image-like tensors (C = 3): all twelve corner values requested before the first use, same arithmetic
        float v[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          v[c][0] = __ldg(pnw + (int64_t)c * HW);
          v[c][1] = __ldg(pne + (int64_t)c * HW);
          v[c][2] = __ldg(psw + (int64_t)c * HW);
          v[c][3] = __ldg(pse + (int64_t)c * HW);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float vnw = g.oknw ? v[c][0] : 0.f, vne = g.okne ? v[c][1] : 0.f;
          const float vsw = g.oksw ? v[c][2] : 0.f, vse = g.okse ? v[c][3] : 0.f;
          float acc = vnw * g.wnw;
          acc = fmaf(vne, g.wne, acc);
          acc = fmaf(vsw, g.wsw, acc);
          acc = fmaf(vse, g.wse, acc);
          st_stream(oc + (int64_t)c * HW, HAS_MASK ? __fmul_rn(acc, m) : acc);
        }
      } else {
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
    }
    if (USE_TMA) {
      __syncthreads();  // every thread has consumed buffer `buf` before it is refilled
      buf ^= 1;
    }
  }
}

// The row loop of fwd_nhwc_kernel: LP lanes at a time stream a pixel's float4 channel groups.
template <int LP, int QI, bool HAS_MASK, bool OTHER>
__device__ __forceinline__ void fwd_nhwc_rows(const uint4* __restrict__ off_row, const float4* __restrict__ w_row,
                                              const float2* __restrict__ mk_row, const char* xl, char* ol,
                                              int64_t ot_delta, int npx, int grp, int nq, uint32_t pxb) {
  constexpr int G = 32 / LP;
#pragma unroll 1
  for (int s = 0; s < npx; s += G) {
    const int pa = s + grp;
    if (G == 1 || pa < npx) {
      const uint4 off = off_row[pa];
      const float4 w = w_row[pa];
      const float2 mk = mk_row[pa];
      const int ok = __float_as_int(mk.y);
      const char* px = xl;
      char* po = ol;
#pragma unroll 1
      for (int qi = 0; qi < (QI > 0 ? QI : nq); ++qi) {
        float4 a = ldg_batch(reinterpret_cast<const float4*>(px + off.x));
        float4 b = ldg_batch(reinterpret_cast<const float4*>(px + off.y));
        float4 c = ldg_batch(reinterpret_cast<const float4*>(px + off.z));
        float4 e = ldg_batch(reinterpret_cast<const float4*>(px + off.w));
        float4 ot = make_float4(0.f, 0.f, 0.f, 0.f);
        if (OTHER) ot = ldg_batch(reinterpret_cast<const float4*>(po + ot_delta));
        if (ok != 15) {  // a corner outside the image (zeros padding / exact border hits)
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(ok & 1)) a = z;
          if (!(ok & 2)) b = z;
          if (!(ok & 4)) c = z;
          if (!(ok & 8)) e = z;
        }
        float4 o;
        o.x = fmaf(e.x, w.w, fmaf(c.x, w.z, fmaf(b.x, w.y, a.x * w.x)));
        o.y = fmaf(e.y, w.w, fmaf(c.y, w.z, fmaf(b.y, w.y, a.y * w.x)));
        o.z = fmaf(e.z, w.w, fmaf(c.z, w.z, fmaf(b.z, w.y, a.z * w.x)));
        o.w = fmaf(e.w, w.w, fmaf(c.w, w.z, fmaf(b.w, w.y, a.w * w.x)));
        if (HAS_MASK) {
          o.x = __fmul_rn(o.x, mk.x);
          o.y = __fmul_rn(o.y, mk.x);
          o.z = __fmul_rn(o.z, mk.x);
          o.w = __fmul_rn(o.w, mk.x);
          if (OTHER) {  // out = m*warp + (1-m)*other
            const float om = 1.f - mk.x;
            o.x = __fadd_rn(o.x, __fmul_rn(om, ot.x));
            o.y = __fadd_rn(o.y, __fmul_rn(om, ot.y));
            o.z = __fadd_rn(o.z, __fmul_rn(om, ot.z));
            o.w = __fadd_rn(o.w, __fmul_rn(om, ot.w));
          }
        }
        st_stream(reinterpret_cast<float4*>(po), o);
        px += LP * 16;
        po += LP * 16;
      }
    }
    ol += G * pxb;
  }
}

// ---------------------------------------------------------------------------------------------
// channels-last.  One CTA = one 8 x 32 pixel tile (non-persistent grid: the block scheduler keeps the
// SMs full); its flow/mask planes arrive by TMA.  After the one mbarrier wait the eight warps never
// synchronise with each other again: warp w owns tile row w, every lane computes the geometry of one
// pixel into the warp's private shared-memory slice (corner positions as 32-bit BYTE offsets into the
// image, so that an address is one add), then LP lanes at a time stream a pixel's float4 channel
// groups: four 128-bit corner loads issued back to back, one 128-bit streaming store.
//   LP  lanes per pixel (a warp moves 32/LP pixels side by side)
//   QI  float4 groups per lane when C/4 == LP*QI exactly, 0 = run-time channel loop
#ifndef C2M_FWD_CTAS
#define C2M_FWD_CTAS 6  // resident CTAs per SM the register budget is set for (tuning switch, tools/build_variants.py)
#endif
template <int LP, int QI, bool HAS_MASK, bool USE_TMA>
__global__ void __launch_bounds__(256, C2M_FWD_CTAS) fwd_nhwc_kernel(const __grid_constant__ FwdParams p,
                                                          const __grid_constant__ CUtensorMap tm_flow,
                                                          const __grid_constant__ CUtensorMap tm_mask) {
  constexpr int TH = 8, TW = 32;
  constexpr int G = 32 / LP;  // pixels a warp moves side by side
  __shared__ alignas(128) float s_flow[2][TH][TW];
  __shared__ alignas(128) float s_mask[TH][TW];
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint4 s_off[TH][TW];   // byte offsets of nw, ne, sw, se inside the image
  __shared__ float4 s_w[TH][TW];    // bilinear weights
  __shared__ float2 s_mk[TH][TW];   // mask value, in-bounds bits
  const Dims& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int t = blockIdx.x;
  const int bx = t % tiles_x;
  const int r = t / tiles_x;
  const int by = r % tiles_y;
  const int n = r / tiles_y;
  const int HW = d.H * d.W;
  const int i = by * TH + warp, j = bx * TW + lane;
  const bool live = (i < d.H) & (j < d.W);
  if (p.pf_tiles >= 0 && lane == 0) {
    // pull the x rows that lie under a tile `pf_tiles` ahead into L2 (its footprint is that region
    // shifted by the flow): by the time that tile runs, its gathers are L2 hits
    const int tp = t + p.pf_tiles;
    const int pbx = tp % tiles_x, pr = tp / tiles_x, pby = pr % tiles_y, pn = pr / tiles_y;
    const int pi = pby * TH + warp;
    if (pn < d.N && pi < d.H) {
      const uint32_t cb = (uint32_t)d.C * 4u;
      prefetch_l2(reinterpret_cast<const char*>(p.x) + ((int64_t)(pn % d.x_batch) * HW + (int64_t)pi * d.W + pbx * TW) * cb,
                  (uint32_t)min(TW, d.W - pbx * TW) * cb);
    }
  }
  float fx = 0.f, fy = 0.f, m = 1.f;
  if (USE_TMA) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
      constexpr uint32_t bytes = (HAS_MASK ? 3u : 2u) * TH * TW * sizeof(float);
      mbar_expect_tx(&bar, bytes);
      tma_load_3d(&s_flow[0][0][0], &tm_flow, &bar, bx * TW, by * TH, n * 2);
      if (HAS_MASK) tma_load_3d(&s_mask[0][0], &tm_mask, &bar, bx * TW, by * TH, n);
    }
    mbar_wait(&bar, 0);
    fx = s_flow[0][warp][lane];
    fy = s_flow[1][warp][lane];
    if (HAS_MASK) m = s_mask[warp][lane];
  } else if (live) {
    fetch_flow_mask(d, p.flow, HAS_MASK ? p.mask : nullptr, n, i, j, fx, fy, m);
  }
  if (i >= d.H) return;  // whole warp
  const uint32_t pxb = (uint32_t)d.C * 4u;  // bytes per pixel
  {
    Geo g;
    make_geo<false>(d, fx, fy, i, min(j, d.W - 1), g);
    s_off[warp][lane] = make_uint4((uint32_t)(g.y0 * d.W + g.x0) * pxb, (uint32_t)(g.y0 * d.W + g.x1) * pxb,
                                   (uint32_t)(g.y1 * d.W + g.x0) * pxb, (uint32_t)(g.y1 * d.W + g.x1) * pxb);
    s_w[warp][lane] = make_float4(g.wnw, g.wne, g.wsw, g.wse);
    const int ok = (int)g.oknw | ((int)g.okne << 1) | ((int)g.oksw << 2) | ((int)g.okse << 3);
    s_mk[warp][lane] = make_float2(m, __int_as_float(ok));
  }
  __syncwarp();
  const int lq = lane % LP, grp = lane / LP;
  const int npx = min(TW, d.W - bx * TW);  // live pixels of this row segment (warp-uniform)
  // blockIdx.y: channel slice of p.cchunk channels (small levels only: more CTAs than tiles)
  const int C4 = p.cchunk >> 2;
  const uint32_t cb0 = blockIdx.y * (uint32_t)p.cchunk * 4u + lq * 16;
  const char* xl = reinterpret_cast<const char*>(p.x) + (int64_t)(n % d.x_batch) * HW * pxb + cb0;
  char* ol = reinterpret_cast<char*>(p.out) + ((int64_t)n * HW + (int64_t)i * d.W + bx * TW + grp) * pxb + cb0;
  // blend operand (out = m*warp + (1-m)*other): the pixel's own row of `other`, same layout as `out`
  const int64_t ot_delta = (HAS_MASK && p.other) ? reinterpret_cast<const char*>(p.other) - reinterpret_cast<const char*>(p.out) : 0;
  const int nq = QI > 0 ? QI : (C4 - lq + LP - 1) / LP;  // float4 groups of this lane
  // (the blend operand gets its own copy of the row loop: the plain one stays as lean as it was)
  if (ot_delta != 0)
    fwd_nhwc_rows<LP, QI, HAS_MASK, true>(s_off[warp], s_w[warp], s_mk[warp], xl, ol, ot_delta, npx, grp, nq, pxb);
  else
    fwd_nhwc_rows<LP, QI, HAS_MASK, false>(s_off[warp], s_w[warp], s_mk[warp], xl, ol, 0, npx, grp, nq, pxb);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
TileMaps make_tile_maps(const Dims& d, const float* flow, const float* mask, int TH, int TW) {
  TileMaps m;
  memset(&m, 0, sizeof(m));
  // (a flow / mask that is resized on the fly has no tile to fetch: plain loads through fetch_flow_mask)
  m.ok = !(d.flags & C2M_FLAG_NO_TMA) && !d.rs.on &&
         make_tensor_map_3d(&m.flow, flow, d.W, d.H, (int64_t)d.N * 2, TW, TH, 2);
  if (m.ok && mask) m.ok = make_tensor_map_3d(&m.mask, mask, d.W, d.H, d.N, TW, TH, 1);
  return m;
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
  const TileMaps tm = make_tile_maps(d, p.flow, p.mask, TH, TW);
  constexpr int NT = TH * TW;
#define C2M_LAUNCH(MASK, TMA)                                                                  \
  do {                                                                                            \
    auto kfn = fwd_nchw_kernel<TH, TW, MASK, TMA, UNROLL>;                                     \
    int cap = resident_ctas(reinterpret_cast<const void*>(kfn), NT) / ysplit;                  \
    if (cap < 1) cap = 1;                                                                      \
    kfn<<<dim3(tiles < cap ? tiles : cap, ysplit), NT, 0, st>>>(p, tm.flow, tm.mask);          \
  } while (0)
  if (p.mask) {
    if (tm.ok) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tm.ok) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
  return C2M_OK;
}

template <int LP, int QI>
static int launch_nhwc_t(FwdParams p, cudaStream_t st) {
  p.pf_tiles = prefetch_tiles(-1);  // measured: the forward gains nothing from it
  constexpr int TH = 8, TW = 32;
  const Dims& d = p.d;
  const int tiles = d.N * ((d.H + TH - 1) / TH) * ((d.W + TW - 1) / TW);
  const TileMaps tm = make_tile_maps(d, p.flow, p.mask, TH, TW);
#define C2M_LAUNCH(MASK, TMA) \
  fwd_nhwc_kernel<LP, QI, MASK, TMA><<<dim3(tiles, d.C / p.cchunk), TH * TW, 0, st>>>(p, tm.flow, tm.mask)
  if (p.mask) {
    if (tm.ok) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tm.ok) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
  return C2M_OK;
}

// C/4 float4 groups per pixel -> (lanes per pixel, groups per lane)
static int launch_nhwc(FwdParams p, cudaStream_t st) {
  // small pyramid levels have fewer tiles than the machine has CTA slots: slice the channels over
  // blockIdx.y (each slice recomputes the tile geometry; slices stay whole 256-byte rows)
  const Dims& d = p.d;
  const int C4 = d.C / 4 / channel_slices(d.N, d.C, d.H, d.W);
  p.cchunk = C4 * 4;
  switch (C4) {
    case 1: return launch_nhwc_t<1, 1>(p, st);
    case 2: return launch_nhwc_t<2, 1>(p, st);
    case 4: return launch_nhwc_t<4, 1>(p, st);
    case 8: return launch_nhwc_t<8, 1>(p, st);
    case 16: return launch_nhwc_t<16, 1>(p, st);
    case 32: return launch_nhwc_t<16, 2>(p, st);
    case 64: return launch_nhwc_t<16, 4>(p, st);
    case 128: return launch_nhwc_t<32, 4>(p, st);
    default: break;
  }
  if (C4 >= 24) return launch_nhwc_t<32, 0>(p, st);
  if (C4 >= 12) return launch_nhwc_t<16, 0>(p, st);
  if (C4 >= 6) return launch_nhwc_t<8, 0>(p, st);
  return launch_nhwc_t<4, 0>(p, st);
}

static int launch_fwd_impl(const FwdParams& p, Layout lx, Layout lo, cudaStream_t st);

int launch_fwd(const FwdParams& p, Layout lx, Layout lo, cudaStream_t st) {
  profile_begin(st);  // the forward is a single kernel
  const int rc = launch_fwd_impl(p, lx, lo, st);
  profile_end(st);
  return rc;
}

static int launch_fwd_impl(const FwdParams& p, Layout lx, Layout lo, cudaStream_t st) {
  const Dims& d = p.d;
  const bool generic = (d.flags & (C2M_FLAG_FORCE_GENERIC | C2M_FLAG_COORD_GRID | C2M_FLAG_TRUE_DIV | C2M_FLAG_NO_FMA)) ||
                       lx != lo || lx == LAYOUT_OTHER || (int64_t)d.H * d.W >= (1ll << 30);
  if (!generic && lx == LAYOUT_NCHW) {
    const int variant = (d.flags >> 24) & 0xf;  // tuning hook (bench sweeps); 0 = default
    switch (variant) {
      case 1: return launch_nchw_t<4, 64, 8>(p, st);
      case 2: return launch_nchw_t<8, 32, 8>(p, st);
      case 3: return launch_nchw_t<16, 32, 8>(p, st);
      case 4: return launch_nchw_t<8, 32, 4>(p, st);
      case 5: return launch_nchw_t<4, 32, 8>(p, st);
      default: return launch_nchw_t<8, 64, 8>(p, st);
    }
  }
  // the channels-last kernel addresses corners by 32-bit byte offsets inside one image
  if (!generic && lx == LAYOUT_NHWC && (d.C % 4) == 0 && ((uintptr_t)p.x % 16) == 0 && ((uintptr_t)p.out % 16) == 0 &&
      ((uintptr_t)p.other % 16) == 0 &&
      (int64_t)d.H * d.W * d.C < (1ll << 30))
    return launch_nhwc(p, st);
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
