// warp_bwd_gather.cu -- gather-form backward of the fused warp + occlusion blend.
//
// grad-input of a bilinear warp is a scatter (ATen: zero-fill + 4 atomicAdd per element).  Here it
// is turned into a gather so that grad-input is written exactly once with plain coalesced stores.
//
// channels-last (the fast path; also NCHW tensors staged through channels-last copies, warp_bwd.cu):
//   segbin_kernel      one warp per row segment of 32 output pixels: geometry once per pixel (kept as a
//                      16-byte record), segment registered as a candidate with the 8 x 32 destination tiles
//                      its samples touch (one global atomic per tile, not per contribution).
//   gather_nhwc_kernel one CTA per tile: local binning (candidate segments -> per-pixel contributor lists
//                      in shared memory), then every pixel in two roles: destination (grad-input =
//                      sum over its list of w * gout[src]) and output (grad-flow / grad-mask from gout and
//                      the four corners of x, reduced over channels).  DET flavour: integer accumulation.
//   overflow_kernel    contributions that went through no list (list full, incoherent segments).
// NCHW kernels (C2M_FLAG_NO_STAGE, or C % 4 != 0):
//   bin_kernel         one thread per output pixel appends (source, weight*mask) to global per-destination
//                      lists (kListCap slots, 32-bit global atomic per claim).
//   gather_nchw_kernel one thread per pixel, channel loop, same two roles.
//
// Algorithmic bytes per pixel: read gout (C) + read x (C) + write gx (C) + flow/mask in, gflow/gmask
// out.  Measured DRAM traffic of the channels-last pipeline: 1.02 - 1.05 x that (bench.py roofline.traffic); no
// zero-fill, no read-modify-write of grad-input.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace c2m {

// ---------------------------------------------------------------------------------------------
// Contributor lists live in the workspace as kListCap/2 planes of int4 -- plane k holds entries 2k and
// 2k+1 (source pixel, weight, source pixel, weight) of every destination pixel -- so that a warp
// reading the lists of 32 neighbouring destinations moves whole 512-byte segments.
__device__ __forceinline__ ListEntry* entry_slot(void* entries, int64_t ndest, int64_t D, int slot) {
  return reinterpret_cast<ListEntry*>(entries) + (((int64_t)(slot >> 1) * ndest + D) << 1) + (slot & 1);
}

// Appends output pixel `idx` = (n, i, j) to the lists of its up-to-four destination pixels.  Returns
// the corners whose destination list was full (bit k), also recorded in p.ovf[idx].
__device__ __forceinline__ unsigned bin_pixel(const BwdParams& p, int64_t idx, int n, int i, int j, float fx, float fy,
                                              float m) {
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const int64_t ndest = (int64_t)HW * d.x_batch;
  Geo g;
  make_geo<true>(d, fx, fy, i, j, g);
  const int dbase = (n % d.x_batch) * HW;
  const int D[4] = {dbase + g.y0 * d.W + g.x0, dbase + g.y0 * d.W + g.x1, dbase + g.y1 * d.W + g.x0,
                    dbase + g.y1 * d.W + g.x1};
  const float ws[4] = {g.wnw * m, g.wne * m, g.wsw * m, g.wse * m};
  const bool act[4] = {g.oknw && ws[0] != 0.f, g.okne && ws[1] != 0.f, g.oksw && ws[2] != 0.f,
                       g.okse && ws[3] != 0.f};
  int slot[4];
  // the four slot claims are independent: issue them back to back (one L2 round trip, not four)
#pragma unroll
  for (int k = 0; k < 4; ++k) slot[k] = act[k] ? atomicAdd(p.cnt + D[k], 1) : 0;
  unsigned ovf = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (act[k]) {
      if (slot[k] < kListCap) {
        // one 8-byte store: (source key, weight)
        *reinterpret_cast<int2*>(entry_slot(p.entries, ndest, D[k], slot[k])) =
            make_int2((int)((uint32_t)idx * (uint32_t)p.key_mul), __float_as_int(ws[k]));
      } else {
        ovf |= 1u << k;
      }
    }
  }
  if (ovf) p.ovf[idx] = (unsigned char)ovf;
  return ovf;
}

constexpr int kBinPixelsPerBlock = 2048;

__global__ void __launch_bounds__(256) bin_kernel(const BwdParams p) {
  __shared__ int s_list[kBinPixelsPerBlock];  // output pixels of this block with a list overflow
  __shared__ int s_n, s_base;
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const int64_t total = (int64_t)HW * (p.n0 + p.nframes);
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int64_t blk0 = (int64_t)HW * p.n0 + (int64_t)blockIdx.x * kBinPixelsPerBlock;
#pragma unroll 1
  for (int it = 0; it < kBinPixelsPerBlock / 256; ++it) {
    const int64_t idx = blk0 + it * 256 + threadIdx.x;
    if (idx >= total) break;
    const int n = (int)(idx / HW);
    const int r = (int)(idx - (int64_t)n * HW);
    const int i = r / d.W, j = r - i * d.W;
    const float fx = __ldg(p.flow + (int64_t)n * 2 * HW + r);
    const float fy = __ldg(p.flow + (int64_t)n * 2 * HW + HW + r);
    const float m = p.mask ? __ldg(p.mask + idx) : 1.f;
    if (bin_pixel(p, idx, n, i, j, fx, fy, m)) s_list[atomicAdd(&s_n, 1)] = (int)idx;
  }
  __syncthreads();
  const int nl = s_n;
  if (nl == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(p.ovf_count, nl);
  __syncthreads();
  for (int k = threadIdx.x; k < nl; k += 256) p.ovf_list[s_base + k] = s_list[k];
}

// ---------------------------------------------------------------------------------------------
// Channels-last local binning, pass A.  One warp per row segment (32 consecutive output pixels of one image
// row): the bounding box of the destination pixels its samples touch is turned into the set of 8 x 32
// destination tiles it overlaps, and the segment registers itself as a candidate with each of them (one
// global atomic per overlapped tile -- about 3 per 32 pixels instead of 4 per pixel).  The gather CTA of a
// tile later recomputes the geometry of its candidates and builds the per-pixel lists in shared memory.
// A segment whose box spans too many tiles (incoherent flow), or whose registration does not fit a tile's
// candidate array, hands the affected contributions to overflow_kernel instead (flag byte + compact list).
__device__ __forceinline__ const float4* key_ptr(const char* base, int key) {
  return reinterpret_cast<const float4*>(base + ((int64_t)(uint32_t)key << 4));  // keys count 16-byte units
}

// The destination tiles (8 x 32, numbered inside one image) that a pixel's in-image, non-zero-weight corners fall in:
// at most four, without repeats.  Evaluated from the pixel RECORD (unclamped nw corner, fractions, mask) with the
// weight expressions of the local binning, so that the count pass (segbin_kernel), the fill pass (flex_fill_kernel)
// and the gather (gather_flex_kernel) agree exactly on what a pixel contributes where.
__device__ __forceinline__ int rec_tiles(const Dims& d, int ux0, int uy0, float ax, float ay, float sm, int tiles_x,
                                         int (&tl)[4]) {
  const float bxw = 1.f - ax, byw = 1.f - ay;
  const float ws[4] = {(bxw * byw) * sm, (ax * byw) * sm, (bxw * ay) * sm, (ax * ay) * sm};
  int nt = 0;
  tl[0] = tl[1] = tl[2] = tl[3] = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int gx = ux0 + (k & 1), gy = uy0 + (k >> 1);
    if ((unsigned)gx < (unsigned)d.W && (unsigned)gy < (unsigned)d.H && ws[k] != 0.f) {
      const int t = (gy >> 3) * tiles_x + (gx >> 5);
      bool seen = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) seen |= (q < nt) && (tl[q] == t);
      if (!seen) tl[nt++] = t;
    }
  }
  return nt;
}

// one row segment: frame n, segment rs of the frame (row-major over (image row, 32-pixel column block))
//   MAXACC  deterministic mode: the two maxima of the fixed-point scale are formed on the way (p.maxacc)
template <bool MAXACC>
__device__ __forceinline__ void segbin_segment(const BwdParams& p, const int n, const int rs, const int lane) {
  constexpr int TH = 8, TW = 32, kMaxCells = 12;
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int nb = d.x_batch == d.N ? n : n % d.x_batch;      // image of x the frame samples
  // (the kernel is bound by integer instructions: no integer divisions on the common path)
  const int i = d.H * tiles_x < (1 << 20) ? __float2int_rd(((float)rs + 0.5f) * __frcp_rn((float)tiles_x)) : rs / tiles_x;
  const int bx = rs - i * tiles_x;
  const int j = bx * TW + lane;
  const bool live = j < d.W;
  int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
  int ys[4] = {0, 0, 0, 0}, xs[4] = {0, 0, 0, 0};
  bool act[4] = {false, false, false, false};
  unsigned inimg = 0;  // corners inside the image
  const int idx = n * HW + i * d.W + j;
  int r_ux0 = -100, r_uy0 = -100;  // the pixel record, kept for the incoherent route
  float r_ax = 0.f, r_ay = 0.f, r_m = 0.f;
  // deterministic mode, called once on every way out: per destination an upper bound of the contributions its
  // list will see (all in-image corners, whatever their weight), or bit 30 for a contribution that goes to
  // overflow_kernel instead; zero_hot_rows_kernel clears the accumulator rows of the destinations that can
  // receive terms outside their 12-entry list (a per-tile mark written with plain stores instead of the
  // bit-30 atomics was tried: the reads of the small mark array hot-spot L2 and it is slower)
  bool all_hot = false;  // this image already has so many incoherent segments that all its rows get cleared
  // (warp-aggregating these atomics with __match_any_sync was measured: it takes 7 % off the kernel for clamped
  // out-of-bounds flows and adds 70 % for coherent ones)
  auto tally = [&](unsigned ovfbits) {
    if (!p.cnt) return;
    int* c0 = p.cnt + nb * HW;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (inimg & (1u << k)) {
        if (ovfbits & (1u << k)) {
          if (!all_hot) atomicOr(c0 + ys[k] * d.W + xs[k], 0x40000000);  // (counts stay below 2^30)
        } else {
          atomicAdd(c0 + ys[k] * d.W + xs[k], 1);
        }
      }
  };
  float amask = 0.f;  // |mask| of this pixel (deterministic mode: max|mask| is one factor of the fixed-point scale)
  if (MAXACC) {
    // deterministic mode: max|gout| (the other factor) over the segment's own 32 rows rides along -- a streaming read
    // next to this kernel's integer work instead of a separate 0.2 ms pass over gout.  A non-finite value reports
    // +inf (fmaxf drops a NaN operand: test the sum, which is NaN as soon as one element is).
    float mx = 0.f;
    const int n4 = min(TW, d.W - bx * TW) * (d.C >> 2);
    const float4* row = reinterpret_cast<const float4*>(p.gout + ((int64_t)n * HW + i * d.W + bx * TW) * d.C);
#pragma unroll 4
    for (int k = lane; k < n4; k += 32) {
      const float4 v = __ldg(row + k);
      const float m4 = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
      const float chk = v.x + v.y + v.z + v.w;
      mx = (chk == chk) ? fmaxf(mx, m4) : __int_as_float(0x7f800000);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // (most warps find the running maximum already at least as large: one load instead of a same-address atomic)
    if (lane == 0 && mx > 0.f && __float_as_uint(mx) > *reinterpret_cast<volatile unsigned*>(p.maxacc))
      atomicMax(p.maxacc, __float_as_uint(mx));
  }
  if (live) {
    const float* fl = p.flow + (int64_t)n * 2 * HW + i * d.W + j;
    const float fx = __ldg(fl), fy = __ldg(fl + HW);
    const float m = p.mask ? __ldg(p.mask + (int64_t)n * HW + i * d.W + j) : 1.f;
    amask = (m == m) ? fabsf(m) : __int_as_float(0x7f800000);
    Geo g;
    make_geo<true>(d, fx, fy, i, j, g);
    ys[0] = ys[1] = g.y0; ys[2] = ys[3] = g.y1;
    xs[0] = xs[2] = g.x0; xs[1] = xs[3] = g.x1;
    // unclamped corner (-1 and W / H mark out-of-image neighbours; zeros-padding outliers sit at -100)
    const int ux0 = g.oknw | g.oksw ? g.x0 : (g.okne | g.okse ? g.x1 - 1 : -100);
    const int uy0 = g.oknw | g.okne ? g.y0 : (g.oksw | g.okse ? g.y1 - 1 : -100);
    p.pixrec[idx] = make_int4((ux0 & 0xffff) | (uy0 << 16), __float_as_int(g.ax), __float_as_int(g.ay),
                              __float_as_int(m));
    r_ux0 = ux0; r_uy0 = uy0; r_ax = g.ax; r_ay = g.ay; r_m = m;
    act[0] = g.oknw && g.wnw * m != 0.f;
    act[1] = g.okne && g.wne * m != 0.f;
    act[2] = g.oksw && g.wsw * m != 0.f;
    act[3] = g.okse && g.wse * m != 0.f;
    inimg = (unsigned)g.oknw | ((unsigned)g.okne << 1) | ((unsigned)g.oksw << 2) | ((unsigned)g.okse << 3);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (act[k]) {
        xmin = min(xmin, xs[k]); xmax = max(xmax, xs[k]);
        ymin = min(ymin, ys[k]); ymax = max(ymax, ys[k]);
      }
  }
  if (MAXACC && p.mask) {
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(amask));  // (non-negative floats order as integers)
    if (lane == 0 && mb > *reinterpret_cast<volatile unsigned*>(p.maxacc + 1)) atomicMax(p.maxacc + 1, mb);
  }
  xmin = __reduce_min_sync(0xffffffffu, xmin);
  xmax = __reduce_max_sync(0xffffffffu, xmax);
  ymin = __reduce_min_sync(0xffffffffu, ymin);
  ymax = __reduce_max_sync(0xffffffffu, ymax);
  if (xmax < xmin) {  // nothing lands anywhere (all weights zero / out of bounds)
    tally(0u);
    return;
  }
  const int tx0 = xmin >> 5, ty0 = ymin >> 3;  // (TW = 32, TH = 8; clamped corners are never negative)
  const int ncols = (xmax >> 5) - tx0 + 1, nrows = (ymax >> 3) - ty0 + 1;
  const int ncell = ncols * nrows;
  unsigned fail;
  if (ncell > kMaxCells && p.bcount) {
    // incoherent flow: the segment's pixels register one by one with the destination tiles they touch (count pass of
    // a counting sort keyed by tile; flex_scan_kernel / flex_fill_kernel complete it, gather_flex_kernel forms those
    // tiles' grad-input) -- no per-contribution atomics on grad-input, in deterministic mode none on the accumulator
    int tl[4] = {-1, -1, -1, -1};
    if (live) rec_tiles(d, r_ux0, r_uy0, r_ax, r_ay, r_m, tiles_x, tl);
    int* bc = p.bcount + nb * tiles_y * tiles_x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // clamped out-of-bounds flows send most of a segment to the same border tile: one atomic per distinct tile
      const unsigned peers = __match_any_sync(0xffffffffu, tl[k]);
      if (tl[k] >= 0 && lane == __ffs(peers) - 1) atomicAdd(bc + tl[k], __popc(peers));
    }
    if (lane == 0) p.iseg_list[atomicAdd(p.iseg_count, 1)] = n * (d.H * tiles_x) + rs;
    tally(0u);
    return;
  }
  if (ncell > kMaxCells) {
    fail = 0xffffffffu;
    if (p.cnt) {
      // deterministic mode: incoherent segments are counted per image; once an image has more than
      // p.incoh_thresh of them zero_hot_rows_kernel clears all its rows, and later segments need not flag
      // their destinations one by one (with fully incoherent flows that is four atomics per pixel saved)
      int k = 0;
      if (lane == 0) k = atomicAdd(p.incoh + nb, 1);
      all_hot = __shfl_sync(0xffffffffu, k, 0) >= p.incoh_thresh;
    }
  } else {
    bool ok = true;
    if (lane < ncell) {
      // lane / ncols: lane < 32, ncols <= 12 -- (lane + 0.5) / ncols is never near an integer
      const int cr = __float2int_rd(((float)lane + 0.5f) * __frcp_rn((float)ncols));
      const int dt = (nb * tiles_y + ty0 + cr) * tiles_x + tx0 + (lane - cr * ncols);
      const int slot = atomicAdd(p.tcnt + dt, 1);
      if (slot < p.cand_cap)
        p.tlist[(int64_t)dt * p.cand_cap + slot] = make_int2(n * HW + i * d.W + bx * TW, min(TW, d.W - bx * TW));
      else ok = false;
    }
    fail = __ballot_sync(0xffffffffu, !ok);
  }
  if (fail == 0u) {
    tally(0u);
    return;
  }
  unsigned ovf = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (act[k]) {
      const int cell = ((ys[k] >> 3) - ty0) * ncols + ((xs[k] >> 5) - tx0);
      if (fail == 0xffffffffu || ((fail >> cell) & 1u)) ovf |= 1u << k;
    }
  // each pixel belongs to exactly one segment and no gather CTA runs yet: a plain byte store is race free
  const unsigned has = __ballot_sync(0xffffffffu, ovf != 0u);
  tally(ovf);
  if (has == 0u) return;
  int base = 0;
  if (lane == __ffs(has) - 1) base = atomicAdd(p.ovf_count, __popc(has));
  base = __shfl_sync(0xffffffffu, base, __ffs(has) - 1);
  if (ovf) {
    p.ovf[idx] = (unsigned char)ovf;
    p.ovf_list[base + __popc(has & ((1u << lane) - 1u))] = idx;
  }
}

template <bool MAXACC>
__global__ void __launch_bounds__(256) segbin_kernel(const BwdParams p) {
  const int rs = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // segment within the frame
  if (rs >= p.d.H * ((p.d.W + 31) / 32)) return;
  segbin_segment<MAXACC>(p, blockIdx.y, rs, threadIdx.x & 31);
}

// Deterministic mode: clear the 64-bit accumulator rows of the destinations that can receive terms -- more
// in-image corners than list slots, or a contribution handed to overflow_kernel -- instead of the whole
// accumulator (one int64 per grad-input element).  One warp per 32 destinations.
__global__ void __launch_bounds__(256) zero_hot_rows_kernel(const int* __restrict__ cnt, const int* __restrict__ incoh,
                                                            int incoh_thresh, long long* __restrict__ acc,
                                                            int64_t ndest, int HW, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
  if (w0 >= ndest) return;
  const int64_t D = w0 + lane;
  const bool hot = D < ndest && (cnt[D] > kLocalCap || incoh[D / HW] > incoh_thresh);
  const unsigned m = __ballot_sync(0xffffffffu, hot);
  if (m == 0u) return;
  // the warp's 32 rows are one contiguous block of 32 * C int64: 16-byte stores, lanes side by side, rows
  // that are not hot skipped (C % 4 == 0 on this path)
  int4* blk = reinterpret_cast<int4*>(acc + w0 * C);
  const int per_row = C >> 1;  // int4 per row
  const int nrows = (int)min((int64_t)32, ndest - w0);
  for (int k = lane; k < nrows * per_row; k += 32)
    if ((m >> (k / per_row)) & 1u) blk[k] = make_int4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// The list tail: contributions that did not fit a destination's in-line slots, applied with atomics
// after the gather has written grad-input.  bin_kernel left a compact list of the affected output
// pixels; one warp per listed pixel, lanes across channels (every lane recomputes the geometry).
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One block takes 256 listed pixels at a time.  Phase A: thread t recomputes the geometry of record t
// into shared memory.  Phase B: a group of GS lanes per record moves the channels (VEC4: channels-last
// with C % 4 == 0 and 16-byte aligned tensors, one 128-bit reduction per four channels and corner).
struct OvfRec {
  int64_t a[4];  // element offsets of the four destination pixels in gx
  int64_t g0;    // element offset of the pixel in gout
  float v[4];    // weight * mask, 0 for corners that did fit their list
  int c0, c1;    // channel range (a slice of a channel-sliced gather, else all)
};

// DET: the terms go to the 64-bit fixed-point accumulator instead (deterministic channels-last gather; runs
// BEFORE the gather, which folds the accumulator rows of the marked destinations into its own sums).
template <bool VEC4, bool DET = false>
__global__ void __launch_bounds__(256) overflow_kernel(const BwdParams p) {
  constexpr int GS = VEC4 ? 16 : 32;
  constexpr int NG = 256 / GS;
  __shared__ OvfRec s_rec[256];
  const Dims& d = p.d;
  const int HW = d.H * d.W;
  const int count = *reinterpret_cast<const volatile int*>(p.ovf_count);
  const int tid = threadIdx.x;
  const int gl = tid % GS, grp = tid / GS;
  for (int base = blockIdx.x * 256; base < count; base += gridDim.x * 256) {
    const int nrec = min(256, count - base);
    if (tid < nrec) {
      int idx = __ldg(p.ovf_list + base + tid);
      OvfRec rec;
      rec.c0 = 0;
      rec.c1 = d.C;
      const unsigned char* flags = p.ovf;
      if (VEC4 && p.cchunk != d.C) {  // channel-sliced gather: the entry carries the slice that overflowed
        const int tag = (int)((unsigned)idx >> 24);
        idx &= 0xffffff;
        flags += (int64_t)tag * p.ovf_stride;
        if (tag) {
          rec.c0 = (tag - 1) * p.cchunk;
          rec.c1 = rec.c0 + p.cchunk;
        }
      }
      const unsigned f = flags[idx];
      const int n = idx / HW;
      const int r = idx - n * HW;
      const int i = r / d.W, j = r - i * d.W;
      const float fx = __ldg(p.flow + (int64_t)n * 2 * HW + r);
      const float fy = __ldg(p.flow + (int64_t)n * 2 * HW + HW + r);
      const float m = p.mask ? __ldg(p.mask + idx) : 1.f;
      Geo g;
      make_geo<true>(d, fx, fy, i, j, g);
      const int64_t xbase = (int64_t)(n % d.x_batch) * p.xs[0];
      rec.a[0] = xbase + g.y0 * p.xs[2] + g.x0 * p.xs[3];
      rec.a[1] = xbase + g.y0 * p.xs[2] + g.x1 * p.xs[3];
      rec.a[2] = xbase + g.y1 * p.xs[2] + g.x0 * p.xs[3];
      rec.a[3] = xbase + g.y1 * p.xs[2] + g.x1 * p.xs[3];
      rec.v[0] = (f & 1u) ? g.wnw * m : 0.f;
      rec.v[1] = (f & 2u) ? g.wne * m : 0.f;
      rec.v[2] = (f & 4u) ? g.wsw * m : 0.f;
      rec.v[3] = (f & 8u) ? g.wse * m : 0.f;
      rec.g0 = (int64_t)n * p.gs[0] + i * p.gs[2] + j * p.gs[3];
      s_rec[tid] = rec;
    }
    __syncthreads();
#pragma unroll 2
    for (int k = grp; k < nrec; k += NG) {
      const OvfRec& rec = s_rec[k];
      if (DET) {
        const float scale = fixed_scale_from(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]), p.count_log2);
        for (int c = gl; c < d.C; c += GS) {
          const float go = p.gout[rec.g0 + c];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = rec.v[q];
            if (v != 0.f)
              atomicAdd(reinterpret_cast<unsigned long long*>(p.acc64 + rec.a[q] + c), (unsigned long long)to_fixed(v * go, scale));
          }
        }
        if (gl == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (rec.v[q] != 0.f) p.touched[rec.a[q] / d.C] = 1;
        }
      } else if (VEC4) {
        for (int c = rec.c0 + gl * 4; c < rec.c1; c += GS * 4) {
          const float4 go = ldg_batch(reinterpret_cast<const float4*>(p.gout + rec.g0 + c));
          float* gx = p.gx + c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = rec.v[q];
            if (v != 0.f) red_add_v4(gx + rec.a[q], make_float4(v * go.x, v * go.y, v * go.z, v * go.w));
          }
        }
      } else {
        for (int c = gl; c < d.C; c += GS) {
          const float go = p.gout[rec.g0 + c * p.gs[1]];
          float* gx = p.gx + c * p.xs[1];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = rec.v[q];
            if (v != 0.f) atomicAdd(gx + rec.a[q], v * go);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int kFlexHeavy = 768;            // a tile with more registered pixels than this is "heavy": gather_flex_kernel
constexpr int kFlexLong = 64;              // a destination list longer than this is summed by all warps of the CTA
constexpr int kFlexChunk = 768;            // candidate pixels per chunk
constexpr int kFlexCap = 4 * kFlexChunk;   // contributions per chunk (four corners each), + padding to even starts

// Incoherent flows: the counting sort's middle passes.
// flex_scan_kernel (one block): exclusive scan of the per-tile registration counts -> start offsets into the pool, and
// the compact list of the heavy tiles (count > kFlexHeavy), in tile order.  Exits at once when no segment was incoherent.
__global__ void __launch_bounds__(1024) flex_scan_kernel(const BwdParams p, int ntile) {
  __shared__ int s_a[32], s_b[32];
  if (*reinterpret_cast<const volatile int*>(p.iseg_count) == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (ntile + 1023) / 1024;
  const int t0 = min(tid * per, ntile), t1 = min(t0 + per, ntile);
  int sum = 0, nz = 0;
  for (int t = t0; t < t1; ++t) {
    const int c = p.bcount[t];
    sum += c;
    nz += c > kFlexHeavy;
  }
  int isum = sum, inz = nz;  // inclusive scans over the block
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, isum, o), b = __shfl_up_sync(0xffffffffu, inz, o);
    if (lane >= o) { isum += a; inz += b; }
  }
  if (lane == 31) { s_a[warp] = isum; s_b[warp] = inz; }
  __syncthreads();
  if (warp == 0) {
    int a = s_a[lane], b = s_b[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ua = __shfl_up_sync(0xffffffffu, a, o), ub = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ua; b += ub; }
    }
    s_a[lane] = a;
    s_b[lane] = b;
  }
  __syncthreads();
  int start = isum - sum + (warp ? s_a[warp - 1] : 0), fpos = inz - nz + (warp ? s_b[warp - 1] : 0);
  for (int t = t0; t < t1; ++t) {
    const int c = p.bcount[t];
    p.bstart[t] = start;
    start += c;
    if (c > kFlexHeavy) p.flex_list[fpos++] = t;
  }
  if (tid == 1023) *p.flex_count = fpos;
}

// flex_fill_kernel: one warp per incoherent segment writes its pixels into the pool ranges of the tiles they touch.
__global__ void __launch_bounds__(256) flex_fill_kernel(const BwdParams p) {
  constexpr int TH = 8, TW = 32;
  const Dims& d = p.d;
  const int nseg = *reinterpret_cast<const volatile int*>(p.iseg_count);
  const int HW = d.H * d.W;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int segs = d.H * tiles_x;
  const int lane = threadIdx.x & 31;
  for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nseg; k += gridDim.x * 8) {
    const int id = __ldg(p.iseg_list + k);
    const int n = id / segs, rs = id - n * segs;
    const int i = rs / tiles_x, bx = rs - i * tiles_x;
    const int j = bx * TW + lane;
    const int sidx = n * HW + i * d.W + j;
    int tl[4] = {-1, -1, -1, -1};
    if (j < d.W) {
      const int4 rec = __ldg(p.pixrec + sidx);
      rec_tiles(d, (int)(short)(rec.x & 0xffff), rec.x >> 16, __int_as_float(rec.y), __int_as_float(rec.z),
                __int_as_float(rec.w), tiles_x, tl);
    }
    const int tb = (n % d.x_batch) * tiles_y * tiles_x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned peers = __match_any_sync(0xffffffffu, tl[q]);
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (tl[q] >= 0 && lane == leader) base = atomicAdd(p.bfill + tb + tl[q], __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (tl[q] >= 0) {
        const int t = tb + tl[q];
        const int pos = base + __popc(peers & ((1u << lane) - 1u));
        if (pos < __ldg(p.bcount + t)) p.pool[__ldg(p.bstart + t) + pos] = sidx;
      }
    }
  }
}

// gather_flex_kernel: grad-input of the flex tiles.  Persistent CTAs take tiles from a work queue.  A tile's
// contributors are the pixels of its candidate segments (coherent neighbours) followed by its pool range; they are
// processed in chunks of kFlexChunk pixels: count per destination pixel (shared-memory atomics), scan, fill -- a
// counting sort by destination inside the tile, lists of any length -- then every destination pixel sums its list
// with LP lanes across the channels (destinations dealt round robin over all lane groups of the CTA: clamped
// out-of-bounds flows pile thousands of contributions on one border row).  Later chunks add to the first one's result.
// DET: order-independent 64-bit fixed-point sums (the pool order is not reproducible).

template <int LP, bool DET>
__global__ void __launch_bounds__(256, 4) gather_flex_kernel(const BwdParams p) {
  constexpr int TH = 8, TW = 32, NPIX = TH * TW;
  __shared__ int s_cnt[NPIX], s_start[NPIX], s_fill[NPIX];
  __shared__ alignas(16) int2 s_e[kFlexCap + NPIX + 4];
  __shared__ int s_tile, s_wsum[8];
  __shared__ union {  // pass 1 of the reduce: per-warp partial sums of one float4 group per lane
    float4 f[8][LP];
    long long i[8][LP][4];
  } s_part;
  const Dims& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int HW = d.H * d.W, C4 = d.C >> 2;
  const int nflex = *reinterpret_cast<const volatile int*>(p.flex_count);
  const int lq = lane % LP;
  const float scale = DET ? fixed_scale_from(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]), p.count_log2) : 1.f;
  const float inv_scale = DET ? fixed_inv_scale(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]), scale) : 1.f;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(p.flex_next, 1);
    __syncthreads();
    const int item = s_tile;
    if (item >= nflex) break;
    const int T = __ldg(p.flex_list + item);
    const int bx = T % tiles_x, r = T / tiles_x, by = r % tiles_y, n = r / tiles_y;  // n: image of x
    const int ncand = min(__ldg(p.tcnt + T), p.cand_cap);
    const int nb = __ldg(p.bcount + T), bs = __ldg(p.bstart + T);
    const int2* tl = p.tlist + (int64_t)T * p.cand_cap;
    const int V = ncand * 32 + nb;  // virtual candidate pixels: the segments' 32 lanes each, then the pool range
    for (int v0 = 0; v0 < V; v0 += kFlexChunk) {
      const bool first = v0 == 0, last = v0 + kFlexChunk >= V;
      s_cnt[tid] = 0;
      s_fill[tid] = 0;
      __syncthreads();
      // ---- pass 1: resolve this thread's candidate pixels, count their contributions per destination
      int sidx[3];
      int4 rec[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int v = v0 + k * 256 + tid;
        sidx[k] = -1;
        if (v < V) {
          if (v < ncand * 32) {
            const int2 seg = __ldg(tl + (v >> 5));
            if ((v & 31) < seg.y) sidx[k] = seg.x + (v & 31);
          } else {
            sidx[k] = __ldg(p.pool + bs + (v - ncand * 32));
          }
        }
        if (sidx[k] >= 0) {
          rec[k] = __ldg(p.pixrec + sidx[k]);
          const int ux0 = (int)(short)(rec[k].x & 0xffff), uy0 = rec[k].x >> 16;
          const int dy = uy0 - by * TH, dx = ux0 - bx * TW;
          if (dy >= -1 && dy < TH && dx >= -1 && dx < TW) {
            const float ax = __int_as_float(rec[k].y), ay = __int_as_float(rec[k].z), sm = __int_as_float(rec[k].w);
            const float bxw = 1.f - ax, byw = 1.f - ay;
            const float ws[4] = {(bxw * byw) * sm, (ax * byw) * sm, (bxw * ay) * sm, (ax * ay) * sm};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int cy = dy + (c >> 1), cx = dx + (c & 1);
              if ((unsigned)cy < (unsigned)TH && (unsigned)cx < (unsigned)TW && uy0 + (c >> 1) < d.H &&
                  ux0 + (c & 1) < d.W && ws[c] != 0.f)
                atomicAdd(&s_cnt[cy * TW + cx], 1);
            }
          } else {
            sidx[k] = -1;
          }
        }
      }
      __syncthreads();
      // ---- exclusive scan of the counts, each rounded up to even (list starts stay 16-byte aligned)
      {
        const int c = (s_cnt[tid] + 1) & ~1;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += u;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        int base = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) base += (w < warp) ? s_wsum[w] : 0;
        s_start[tid] = base + inc - c;
      }
      __syncthreads();
      // ---- pass 2: fill the lists
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (sidx[k] < 0) continue;
        const int ux0 = (int)(short)(rec[k].x & 0xffff), uy0 = rec[k].x >> 16;
        const int dy = uy0 - by * TH, dx = ux0 - bx * TW;
        const float ax = __int_as_float(rec[k].y), ay = __int_as_float(rec[k].z), sm = __int_as_float(rec[k].w);
        const float bxw = 1.f - ax, byw = 1.f - ay;
        const float ws[4] = {(bxw * byw) * sm, (ax * byw) * sm, (bxw * ay) * sm, (ax * ay) * sm};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cy = dy + (c >> 1), cx = dx + (c & 1);
          if ((unsigned)cy < (unsigned)TH && (unsigned)cx < (unsigned)TW && uy0 + (c >> 1) < d.H &&
              ux0 + (c & 1) < d.W && ws[c] != 0.f) {
            const int dp = cy * TW + cx;
            s_e[s_start[dp] + atomicAdd(&s_fill[dp], 1)] = make_int2((int)((uint32_t)sidx[k] * (uint32_t)C4),
                                                                    __float_as_int(ws[c]));
          }
        }
      }
      __syncthreads();
      // ---- reduce: one WARP per destination pixel (the lists are long here): LP lanes across the float4 groups of a
      // row, 32 / LP list entries side by side, four rounds of loads in flight per lane; the side-by-side partial
      // sums meet through shuffles.  (Float: the pool order is not reproducible, neither is this sum -- like the
      // atomics it replaces.  DET: 64-bit fixed point, any order gives the same bits.)
      const char* gl = reinterpret_cast<const char*>(p.gout);
      constexpr int P = 32 / LP;  // entries in flight side by side
      const int sub = lane / LP;
      // a destination with a very long list (the image corners collect thousands of clamped samples) is summed by
      // ALL eight warps, each taking every eighth group of entries; the partial sums meet in shared memory.  Pass 0:
      // the ordinary destinations, one per warp; pass 1: the long ones, one at a time, all warps together.
      for (int pass = 0; pass < 2; ++pass)
      for (int dp = pass ? 0 : warp; dp < NPIX; dp += pass ? 1 : 8) {
        const int i = by * TH + dp / TW, j = bx * TW + dp % TW;
        if (i >= d.H || j >= d.W) continue;
        const int cnt = s_cnt[dp];
        const bool longlist = cnt > kFlexLong;
        if (longlist != (pass == 1)) continue;
        if (cnt == 0 && !first && !(DET && last)) continue;  // nothing to add in this chunk (DET: the last one converts)
        const int2* e = s_e + s_start[dp];
        const int64_t D = (int64_t)n * HW + (int64_t)i * d.W + j;
        const bool fold = DET && first && __ldcg(p.touched + D) != 0;  // terms pushed by overflow_kernel<DET>
        const int kstep = pass ? 8 * 4 * P : 4 * P, koff = pass ? warp * 4 * P : 0;
        for (int q0 = 0; q0 < C4; q0 += LP) {  // (warp-uniform trip count: the shuffles below need every lane)
          const int q = q0 + lq;
          const bool qv = q < C4;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          long long ia[4] = {0, 0, 0, 0};
          for (int k = koff + sub; k < cnt; k += kstep) {
            int2 en[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool on = qv && k + u * P < cnt;
              en[u] = on ? e[k + u * P] : make_int2(0, 0);
              v[u] = ldg_batch_if(key_ptr(gl, en[u].x) + q, on);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float w = __int_as_float(en[u].y);  // 0 when the slot is past the list
              if (DET) {
                ia[0] += to_fixed(w * v[u].x, scale);
                ia[1] += to_fixed(w * v[u].y, scale);
                ia[2] += to_fixed(w * v[u].z, scale);
                ia[3] += to_fixed(w * v[u].w, scale);
              } else {
                acc.x = fmaf(w, v[u].x, acc.x);
                acc.y = fmaf(w, v[u].y, acc.y);
                acc.z = fmaf(w, v[u].z, acc.z);
                acc.w = fmaf(w, v[u].w, acc.w);
              }
            }
          }
#pragma unroll
          for (int o = LP; o < 32; o <<= 1) {  // the P side-by-side partial sums
            if (DET) {
#pragma unroll
              for (int c = 0; c < 4; ++c) ia[c] += __shfl_xor_sync(0xffffffffu, ia[c], o);
            } else {
              acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
              acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
              acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
              acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
            }
          }
          if (pass) {  // the eight warps' partial sums, added in warp order by warp 0
            __syncthreads();
            if (sub == 0 && qv) {
              if (DET) {
#pragma unroll
                for (int c = 0; c < 4; ++c) s_part.i[warp][lq][c] = ia[c];
              } else {
                s_part.f[warp][lq] = acc;
              }
            }
            __syncthreads();
            if (warp != 0) continue;
            if (sub == 0 && qv) {
#pragma unroll
              for (int w = 1; w < 8; ++w) {
                if (DET) {
#pragma unroll
                  for (int c = 0; c < 4; ++c) ia[c] += s_part.i[w][lq][c];
                } else {
                  const float4 o = s_part.f[w][lq];
                  acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                }
              }
            }
          }
          if (sub != 0 || !qv) continue;
          float4* gxp = reinterpret_cast<float4*>(p.gx) + D * C4 + q;
          if (DET) {
            long long* ap = p.acc64 + D * d.C + q * 4;
            if (!first || fold) {
              const longlong2 h0 = __ldcg(reinterpret_cast<const longlong2*>(ap));
              const longlong2 h1 = __ldcg(reinterpret_cast<const longlong2*>(ap + 2));
              ia[0] += h0.x; ia[1] += h0.y; ia[2] += h1.x; ia[3] += h1.y;
            }
            if (last) {
              *gxp = make_float4(__ll2float_rn(ia[0]) * inv_scale, __ll2float_rn(ia[1]) * inv_scale,
                                 __ll2float_rn(ia[2]) * inv_scale, __ll2float_rn(ia[3]) * inv_scale);
            } else {
              *reinterpret_cast<longlong2*>(ap) = make_longlong2(ia[0], ia[1]);
              *reinterpret_cast<longlong2*>(ap + 2) = make_longlong2(ia[2], ia[3]);
            }
          } else {
            if (!first) {
              const float4 o = *gxp;
              acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            *gxp = acc;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// channels-last gather.  One CTA = one 8 x 32 pixel tile, warp w owns tile row w and never waits for
// another warp after the TMA barrier.  Every pixel of the row plays two roles:
//   destination  grad-input[pixel] = sum over its contributor list of w * gout[source]
//   output       grad-flow / grad-mask [pixel] from gout[pixel] and the four corners of x
// Phase 0: lane t prepares pixel t of the row (geometry, list) into the warp's shared-memory slice;
//          unused list slots are filled with (own pixel, weight 0) so that phase 1 needs no predicates
//          for the first four entries.
// Phase 1: LP lanes at a time stream a pixel's float4 channel groups, NQ groups per lane per pass.

// The destination role and the output role of a step are each split into "issue the loads" and "use
// them", so that the fused kernel can put both roles' loads in flight before it waits for either.
template <int LP, int NQ>
__device__ __forceinline__ void gx_issue(const char* gl, int cnt, const int4& e0, const int4& e1, float4 (&v)[4][NQ]) {
  const float4* a0 = key_ptr(gl, e0.x);
  const float4* a1 = key_ptr(gl, e0.z);
  const float4* a2 = key_ptr(gl, e1.x);
  const float4* a3 = key_ptr(gl, e1.z);
  // entries past the count: their loads are predicated off and read as zero.  (Leaving the registers undefined
  // instead saves the clears but makes every destination register live across the whole loop -- a predicated load
  // preserves its old value -- and the kernel spills; unpredicated loads of a dummy row cost L1 bandwidth.  Both
  // measured slower, profiles/r2_gather_variants_ab.txt.)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    v[0][q] = ldg_batch_if(a0 + q * LP, cnt > 0);
    v[1][q] = ldg_batch_if(a1 + q * LP, cnt > 1);
    v[2][q] = ldg_batch_if(a2 + q * LP, cnt > 2);
    v[3][q] = ldg_batch_if(a3 + q * LP, cnt > 3);
  }
}

// `ent` points at pair 0 of this destination's list in shared memory, consecutive pairs `pstride` int4 apart;
// NP pairs in all (pairs 0 and 1 arrive in e0 / e1, already loaded).
template <int LP, int NQ, int NP>
__device__ __forceinline__ void gx_finish(const char* gl, char* po, int cnt, const int4& e0, const int4& e1,
                                          const int4* ent, int pstride, const float4 (&v)[4][NQ], bool store = true) {
  // slots past the count contribute nothing (their loads were predicated off): the accumulation is predicated too
  const float w[4] = {__int_as_float(e0.y), __int_as_float(e0.w), __int_as_float(e1.y), __int_as_float(e1.w)};
  float4 acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (cnt > k) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        acc[q].x = fmaf(w[k], v[k][q].x, acc[q].x);
        acc[q].y = fmaf(w[k], v[k][q].y, acc[q].y);
        acc[q].z = fmaf(w[k], v[k][q].z, acc[q].z);
        acc[q].w = fmaf(w[k], v[k][q].w, acc[q].w);
      }
    }
  }
  if (cnt > 4) {  // long list: two entries at a time
#pragma unroll
    for (int k = 2; k < NP; ++k) {
      if (cnt > 2 * k) {
        const int4 e = ent[k * pstride];
        const float4* b0 = key_ptr(gl, e.x);
        const float4* b1 = key_ptr(gl, e.z);
        const bool two = cnt > 2 * k + 1;
        float4 u[2][NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          u[0][q] = ldg_batch(b0 + q * LP);
          u[1][q] = ldg_batch_if(b1 + q * LP, two);
        }
        const float wa = __int_as_float(e.y), wb = two ? __int_as_float(e.w) : 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          acc[q].x = fmaf(wb, u[1][q].x, fmaf(wa, u[0][q].x, acc[q].x));
          acc[q].y = fmaf(wb, u[1][q].y, fmaf(wa, u[0][q].y, acc[q].y));
          acc[q].z = fmaf(wb, u[1][q].z, fmaf(wa, u[0][q].z, acc[q].z));
          acc[q].w = fmaf(wb, u[1][q].w, fmaf(wa, u[0][q].w, acc[q].w));
        }
      }
    }
  }
  if (store) {  // (clear on "flex" tiles: gather_flex_kernel writes their grad-input)
#pragma unroll
    for (int q = 0; q < NQ; ++q) st_stream(reinterpret_cast<float4*>(po) + q * LP, acc[q]);
  }
}

// Deterministic flavour of gx_finish: integer accumulation (see fixed_scale_from); `hot` destinations
// also have terms in the global accumulator (list overflow, incoherent segments), added before the one
// conversion back to float.
template <int LP, int NQ, int NP>
__device__ __forceinline__ void gx_finish_det(const char* gl, char* po, int cnt, const int4& e0, const int4& e1,
                                              const int4* ent, int pstride, const float4 (&v)[4][NQ], float scale,
                                              float inv_scale, const long long* hot_acc) {
  long long acc[NQ][4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0;
  const float w[4] = {__int_as_float(e0.y), __int_as_float(e0.w), __int_as_float(e1.y), __int_as_float(e1.w)};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (cnt > k) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        acc[q][0] += to_fixed(w[k] * v[k][q].x, scale);
        acc[q][1] += to_fixed(w[k] * v[k][q].y, scale);
        acc[q][2] += to_fixed(w[k] * v[k][q].z, scale);
        acc[q][3] += to_fixed(w[k] * v[k][q].w, scale);
      }
    }
  if (cnt > 4) {
#pragma unroll
    for (int k = 2; k < NP; ++k) {
      if (cnt > 2 * k) {
        const int4 e = ent[k * pstride];
        const float4* b0 = key_ptr(gl, e.x);
        const float4* b1 = key_ptr(gl, e.z);
        const bool two = cnt > 2 * k + 1;
        float4 u[2][NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          u[0][q] = ldg_batch(b0 + q * LP);
          u[1][q] = ldg_batch_if(b1 + q * LP, two);
        }
        const float wa = __int_as_float(e.y), wb = two ? __int_as_float(e.w) : 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          acc[q][0] += to_fixed(wa * u[0][q].x, scale) + to_fixed(wb * u[1][q].x, scale);
          acc[q][1] += to_fixed(wa * u[0][q].y, scale) + to_fixed(wb * u[1][q].y, scale);
          acc[q][2] += to_fixed(wa * u[0][q].z, scale) + to_fixed(wb * u[1][q].z, scale);
          acc[q][3] += to_fixed(wa * u[0][q].w, scale) + to_fixed(wb * u[1][q].w, scale);
        }
      }
    }
  }
  if (hot_acc) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const longlong2 h0 = __ldcg(reinterpret_cast<const longlong2*>(hot_acc + q * LP * 4));
      const longlong2 h1 = __ldcg(reinterpret_cast<const longlong2*>(hot_acc + q * LP * 4 + 2));
      acc[q][0] += h0.x; acc[q][1] += h0.y; acc[q][2] += h1.x; acc[q][3] += h1.y;
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    st_stream(reinterpret_cast<float4*>(po) + q * LP,
              make_float4(__ll2float_rn(acc[q][0]) * inv_scale, __ll2float_rn(acc[q][1]) * inv_scale,
                          __ll2float_rn(acc[q][2]) * inv_scale, __ll2float_rn(acc[q][3]) * inv_scale));
}

template <int NQ>
struct DotRegs {
  float4 a[NQ], b[NQ], c[NQ], e[NQ], g[NQ];
};

template <int LP, int NQ>
__device__ __forceinline__ void dot_issue_x(const char* px, const uint4& off, DotRegs<NQ>& r) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    r.a[q] = ldg_batch(reinterpret_cast<const float4*>(px + off.x) + q * LP);
    r.b[q] = ldg_batch(reinterpret_cast<const float4*>(px + off.y) + q * LP);
    r.c[q] = ldg_batch(reinterpret_cast<const float4*>(px + off.z) + q * LP);
    r.e[q] = ldg_batch(reinterpret_cast<const float4*>(px + off.w) + q * LP);
  }
}
template <int LP, int NQ>
__device__ __forceinline__ void dot_issue_g(const char* pg, DotRegs<NQ>& r) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) r.g[q] = ldg_batch(reinterpret_cast<const float4*>(pg) + q * LP);
}
template <int LP, int NQ>
__device__ __forceinline__ void dot_issue(const char* px, const char* pg, const uint4& off, DotRegs<NQ>& r) {
  dot_issue_x<LP, NQ>(px, off, r);
  dot_issue_g<LP, NQ>(pg, r);
}

// sa..se += sum over this lane's channels of gout[c] * x_corner[c]
template <int NQ>
__device__ __forceinline__ void dot_finish(const DotRegs<NQ>& r, float& sa, float& sb, float& sc, float& se) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    sa = fmaf(r.g[q].x, r.a[q].x, fmaf(r.g[q].y, r.a[q].y, fmaf(r.g[q].z, r.a[q].z, fmaf(r.g[q].w, r.a[q].w, sa))));
    sb = fmaf(r.g[q].x, r.b[q].x, fmaf(r.g[q].y, r.b[q].y, fmaf(r.g[q].z, r.b[q].z, fmaf(r.g[q].w, r.b[q].w, sb))));
    sc = fmaf(r.g[q].x, r.c[q].x, fmaf(r.g[q].y, r.c[q].y, fmaf(r.g[q].z, r.c[q].z, fmaf(r.g[q].w, r.c[q].w, sc))));
    se = fmaf(r.g[q].x, r.e[q].x, fmaf(r.g[q].y, r.e[q].y, fmaf(r.g[q].z, r.e[q].z, fmaf(r.g[q].w, r.e[q].w, se))));
  }
}

// Sum four per-lane partial values over the LP lanes of a pixel with 2 + 1 + log2(LP/4) shuffles instead of
// 4 * log2(LP): after the first exchange a lane carries two of the sums, after the second one.  On return the lanes
// with lq % (LP/4) == 0 hold the complete sum number lq / (LP/4) (0: sa, 1: sb, 2: sc, 3: se).
template <int LP>
__device__ __forceinline__ float reduce4(float sa, float sb, float sc, float se, int lq) {
  static_assert(LP >= 4, "four sums need at least four lanes");
  const bool hi = (lq & (LP / 2)) != 0;
  float r0 = (hi ? sc : sa) + __shfl_xor_sync(0xffffffffu, hi ? sa : sc, LP / 2);
  float r1 = (hi ? se : sb) + __shfl_xor_sync(0xffffffffu, hi ? sb : se, LP / 2);
  const bool hi2 = (lq & (LP / 4)) != 0;
  float r = (hi2 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, hi2 ? r0 : r1, LP / 4);
#pragma unroll
  for (int o = LP / 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

//   LP  lanes per pixel (a warp moves 32/LP pixels side by side)
//   QI  float4 groups per lane when C/4 == LP*QI exactly, 0 = run-time channel loop
// The contributor lists are built by the CTA itself in shared memory from the row segments that segbin_kernel
// registered as candidates for this destination tile (no global lists, no global atomics per contribution).
//   DET     deterministic grad-input: fixed-point integer accumulation
template <int LP, int QI, bool DO_GX, bool DO_GF, bool HAS_MASK, bool USE_TMA, bool DET = false>
__global__ void __launch_bounds__(256, 4) gather_nhwc_kernel(const __grid_constant__ BwdParams p,
                                                             const __grid_constant__ CUtensorMap tm_flow,
                                                             const __grid_constant__ CUtensorMap tm_mask) {
  constexpr int TH = 8, TW = 32;
  constexpr int G = 32 / LP;
  // entries per destination held in shared memory.  Deterministic mode takes 16: a list overflow there is pushed to
  // the global accumulator by the discovering warp (64-bit atomics per channel) while its CTA waits at the barrier
  constexpr int CAP = DET ? kLocalCapDet : kLocalCap;
  constexpr int NP = CAP / 2;                        // as int4 entry pairs
  __shared__ alignas(128) float s_flow[DO_GF ? 2 : 1][TH][TW];
  __shared__ alignas(128) float s_mask[TH][TW];
  __shared__ uint4 s_off[DO_GF ? TH : 1][TW];   // byte offsets of the four corners inside the image
  // ax, ay, mask, flags (bits 0-3: corner inside the image, 4 / 5: x / y coordinate clipped -> zero
  // grad-flow); later the pixel's (gflow_x, gflow_y, gmask)
  __shared__ float4 s_aux[DO_GF ? TH : 1][TW];
  // The mbarrier of the flow / mask tile lives in the first bytes of s_aux, which is first written after every thread
  // has left the wait (the __syncthreads() below).  As a variable of its own it put the CTA 16 bytes over 40 KB: four
  // CTAs then needed the 196 KB shared-memory carve-out and ran with 60 KB of L1 instead of 92 KB (measured: gather
  // 0.950 -> 0.939 ms).
  uint64_t& bar = *reinterpret_cast<uint64_t*>(&s_aux[0][0]);
  // the pixel's four dot products sum_c gout[c] * x_corner[c].  (Deterministic mode, which spends its shared memory on
  // longer lists, parks them in the pixel's corner offsets instead: those are dead once its step has issued its loads.
  // Doing the same here frees 4 KB but costs two spilled registers: measured 0.948 ms against 0.939 ms.)
  __shared__ float4 s_sum[(DO_GF && !DET) ? TH : 1][TW];
  __shared__ int s_cnt[DO_GX ? TH : 1][TW];
  __shared__ int4 s_ent[DO_GX ? NP : 1][DO_GX ? TH : 1][TW];
  const Dims& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int t = blockIdx.x;
  const int bx = t % tiles_x;
  const int r = t / tiles_x;
  const int by = r % tiles_y;
  const int n = p.n0 + r / tiles_y;  // DO_GF: frame; gx-only pass: image of x
  const int HW = d.H * d.W;
  const int C4 = d.C >> 2;
  const int i = by * TH + warp, j = bx * TW + lane;
  const bool live = (i < d.H) & (j < d.W);
  const int pix = i * d.W + j;
  if (p.pf_tiles >= 0 && lane == 0) {
    // pull the gout (and x) rows under a tile `pf_tiles` ahead into L2: that tile's gathers land in
    // this region shifted by the flow
    const int tp = t + p.pf_tiles;
    const int pbx = tp % tiles_x, pr = tp / tiles_x, pby = pr % tiles_y, pn = p.n0 + pr / tiles_y;
    const int pi = pby * TH + warp;
    if (pn < p.n0 + p.nframes && pi < d.H) {
      const uint32_t cb = (uint32_t)d.C * 4u;
      const int64_t px0 = (int64_t)pi * d.W + pbx * TW;
      const uint32_t bytes = (uint32_t)min(TW, d.W - pbx * TW) * cb;
      prefetch_l2(reinterpret_cast<const char*>(p.gout) + ((int64_t)pn * HW + px0) * cb, bytes);
      if (DO_GF) prefetch_l2(reinterpret_cast<const char*>(p.x) + ((int64_t)(pn % d.x_batch) * HW + px0) * cb, bytes);
    }
  }
  float fx = 0.f, fy = 0.f, m = 1.f;
  if (DO_GF) {
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
    } else if (live) {
      const float* fl = p.flow + (int64_t)n * 2 * HW + pix;
      fx = __ldg(fl);
      fy = __ldg(fl + HW);
      if (HAS_MASK) m = __ldg(p.mask + (int64_t)n * HW + pix);
    }
  }
  // incoherent flows register single pixels with the destination tiles they touch (nreg of them here).  A few join the
  // candidate segments in the local binning below; a HEAVY tile (clamped out-of-bounds flows pile thousands on the
  // border tiles) gets its grad-input from gather_flex_kernel and this CTA only plays the output role (CTA-uniform)
  // (both tile counters are requested here, side by side: they head the CTA's critical path)
  // (deterministic mode only -- measured: for the float path the atomics of overflow_kernel are faster than the
  // counting sort, profiles/r2_incoherent_flows.txt; compiled out of the ordinary gather, whose binning loop is the
  // critical path of every CTA)
  constexpr bool POOL = DET;
  const int nreg = (POOL && DO_GX && p.bcount) ? __ldg(p.bcount + (n * tiles_y + by) * tiles_x + bx) : 0;
  const int ncand_all = DO_GX ? __ldg(p.tcnt + (n * tiles_y + by) * tiles_x + bx) : 0;
  const bool do_gx = DO_GX && (!POOL || nreg <= kFlexHeavy);
  // deterministic mode: what decides whether this thread's destination pixel is "hot" (below) is requested now, so that
  // the answers are there when the binning is done (they used to sit on the CTA's critical path)
  int det_dc = 0, det_incoh = 0;
  if (DET && DO_GX && live) {
    det_dc = __ldg(p.cnt + (int64_t)n * HW + pix);
    det_incoh = __ldg(p.incoh + n);
  }
  if (DO_GX) {
    // ---- local binning: every warp walks candidate row segments (32 output pixels each), recomputes their
    // geometry and files the contributions that land inside this tile into the destination's list
    s_cnt[warp][lane] = 0;
    __syncthreads();
    const int T = (n * tiles_y + by) * tiles_x + bx;  // destination tile (n: image of x)
    const int ncand = do_gx ? min(ncand_all, p.cand_cap) : 0;
    const int2* tl = p.tlist + (int64_t)T * p.cand_cap;
    const int2 myid = (warp + 8 * lane < ncand) ? __ldg(tl + warp + 8 * lane) : make_int2(0, 0);  // warp w: c = w, w+8, ...
    const int nit = max((ncand - warp + 7) >> 3, 0);
    // registered pixels of incoherent segments: 32 at a time, warp w takes groups w, w + 8, ...
    const int npool = (POOL && do_gx) ? nreg : 0, pool0 = npool ? __ldg(p.bstart + T) : 0;
    const int nitp = POOL ? max((((npool + 31) >> 5) - warp + 7) >> 3, 0) : 0;
    const float det_scale = DET ? fixed_scale_from(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]),
                                                   p.count_log2) : 1.f;
#pragma unroll 2
    for (int it = 0; it < nit + nitp; ++it) {
      int sidx = -1;
      if (!POOL || it < nit) {
        const int seg0 = __shfl_sync(0xffffffffu, myid.x, it), nlive = __shfl_sync(0xffffffffu, myid.y, it);
        if (lane < nlive) sidx = seg0 + lane;
      } else {
        const int v = (((it - nit) << 3) + warp) * 32 + lane;
        if (v < npool) sidx = __ldg(p.pool + pool0 + v);
      }
      // deterministic mode: contributions that found their list full (bit k), kept for the warp-wide push below
      unsigned failbits = 0;
      int f_pos = 0, f_sidx = 0;
      float f_ax = 0.f, f_ay = 0.f, f_m = 0.f;
      if (sidx >= 0) {
        const int4 rec = __ldg(p.pixrec + sidx);
        const int ux0 = (int)(short)(rec.x & 0xffff), uy0 = rec.x >> 16;
        // tile-local position of the nw corner; the four corners are (dy, dx), (dy, dx+1), (dy+1, dx), (dy+1, dx+1)
        const int dy = uy0 - by * TH, dx = ux0 - bx * TW;
        if (dy >= -1 && dy < TH && dx >= -1 && dx < TW) {  // at least one corner can fall inside this tile
          const float ax = __int_as_float(rec.y), ay = __int_as_float(rec.z), sm = __int_as_float(rec.w);
          // same expressions as make_geo: (x1 - ix) == 1 - ax and (ix - x0) == ax bit for bit
          const float bxw = 1.f - ax, byw = 1.f - ay;
          const float ws[4] = {(bxw * byw) * sm, (ax * byw) * sm, (bxw * ay) * sm, (ax * ay) * sm};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cy = dy + (k >> 1), cx = dx + (k & 1);
            const int gy = uy0 + (k >> 1), gxx = ux0 + (k & 1);
            if ((unsigned)cy < (unsigned)TH && (unsigned)cx < (unsigned)TW && gy < d.H && gxx < d.W && ws[k] != 0.f) {
              const int slot = atomicAdd(&s_cnt[cy][cx], 1);
              if (slot < CAP) {
                reinterpret_cast<int2*>(&s_ent[slot >> 1][cy][cx])[slot & 1] =
                    make_int2((int)((uint32_t)sidx * (uint32_t)C4), __float_as_int(ws[k]));
              } else if (DET) {
                failbits |= 1u << k;
                f_pos = rec.x; f_sidx = sidx; f_ax = ax; f_ay = ay; f_m = sm;
              } else {
                // list full: hand the contribution to overflow_kernel (flag byte per output pixel, shared
                // with other tiles' overflows of the same pixel -> word-wide atomic OR)
                const unsigned tag = gridDim.y > 1 ? blockIdx.y + 1 : 0;  // sliced: this slice's own flags
                unsigned* word = reinterpret_cast<unsigned*>(p.ovf + (int64_t)tag * p.ovf_stride) + (sidx >> 2);
                const int sh = (sidx & 3) * 8;
                const unsigned old = atomicOr(word, (1u << k) << sh);
                if (((old >> sh) & 0xffu) == 0u) p.ovf_list[atomicAdd(p.ovf_count, 1)] = sidx | (int)(tag << 24);
              }
            }
          }
        }
      }
      if (DET) {
        // list overflow in deterministic mode: the whole warp adds the term to the destination's global
        // fixed-point row (lanes across channels) and marks the destination; phase 1 of this same CTA
        // picks the row up before it converts
        unsigned any = __ballot_sync(0xffffffffu, failbits != 0u);
        while (any) {
          const int src = __ffs(any) - 1;
          any &= any - 1;
          const unsigned fb = __shfl_sync(0xffffffffu, failbits, src);
          const int pos = __shfl_sync(0xffffffffu, f_pos, src), sx = __shfl_sync(0xffffffffu, f_sidx, src);
          const float ax = __shfl_sync(0xffffffffu, f_ax, src), ay = __shfl_sync(0xffffffffu, f_ay, src);
          const float sm = __shfl_sync(0xffffffffu, f_m, src);
          const int ux0 = (int)(short)(pos & 0xffff), uy0 = pos >> 16;
          const float bxw = 1.f - ax, byw = 1.f - ay;
          const float ws[4] = {(bxw * byw) * sm, (ax * byw) * sm, (bxw * ay) * sm, (ax * ay) * sm};
          const float* gsrc = p.gout + (int64_t)sx * d.C;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (fb & (1u << k)) {
              const int64_t D = (int64_t)n * HW + (int64_t)(uy0 + (k >> 1)) * d.W + (ux0 + (k & 1));
              for (int c = lane; c < d.C; c += 32)
                atomicAdd(reinterpret_cast<unsigned long long*>(p.acc64 + D * d.C + c),
                          (unsigned long long)to_fixed(ws[k] * __ldg(gsrc + c), det_scale));
              if (lane == 0) p.touched[D] = 1;
            }
        }
      }
    }
    // (deterministic mode: the terms and marks a warp pushed to global memory are only ever read by THIS CTA, after
    // this barrier, with L1-bypassing loads -- __syncthreads() orders them; a device-wide fence here cost an L1
    // invalidation per pushing warp, CCTL.IVALL, and with it the gather's L1 hit rate)
    __syncthreads();
  }
  // (the lists stay where local binning put them: phase 1 clamps the count and ignores the slots past it)
  if (DO_GX && DET && live && do_gx) {
    // Deterministic mode.  A destination whose terms all sit in its shared-memory list (no list overflow, nothing
    // in the global accumulator) is summed in FLOAT over the list SORTED BY SOURCE: the set of (source, weight)
    // entries is the same in every run, only their arrival order is not, so a fixed order gives fixed bits at the
    // speed of the ordinary gather.  The rare "hot" destinations (bit 30) keep the order-independent integer sum.
    // "Hot" is decided from run-independent quantities only -- WHICH contributions end up in the accumulator depends
    // on the order of atomic slot claims, and a destination must take the same route in every run: more in-image
    // corners point at it than a list holds (segbin_kernel's tally), a contribution was handed to overflow_kernel
    // (tally bit 30: incoherent segments; failed registrations only happen in over-full tiles), its tile's candidate
    // array is over-full, or its image has more incoherent segments than the clear-everything threshold.
    const int c = s_cnt[warp][lane];
    if (det_dc > CAP || ncand_all > p.cand_cap || det_incoh > p.incoh_thresh) {
      // bit 29: the destination also has terms in the global accumulator row (that row was cleared for it)
      s_cnt[warp][lane] = c | 0x40000000 | (__ldcg(p.touched + (int64_t)n * HW + pix) ? 0x20000000 : 0);
    } else {
      int2* const e = reinterpret_cast<int2*>(&s_ent[0][warp][lane]);  // slot k at e[(k >> 1) * (TH * TW * 2) + (k & 1)]
      for (int a = 1; a < c; ++a) {  // insertion sort, 4 entries on average
        const int2 key = e[(a >> 1) * (TH * TW * 2) + (a & 1)];
        int b = a - 1;
        while (b >= 0) {
          const int2 o = e[(b >> 1) * (TH * TW * 2) + (b & 1)];
          if ((unsigned)o.x <= (unsigned)key.x) break;
          e[((b + 1) >> 1) * (TH * TW * 2) + ((b + 1) & 1)] = o;
          --b;
        }
        e[((b + 1) >> 1) * (TH * TW * 2) + ((b + 1) & 1)] = key;
      }
    }
  }
  if (DO_GF && USE_TMA) {
    mbar_wait(&bar, 0);
    __syncthreads();  // nobody polls the barrier any more when s_aux (whose first bytes it occupies) is written
    fx = s_flow[0][warp][lane];
    fy = s_flow[1][warp][lane];
    if (HAS_MASK) m = s_mask[warp][lane];
  }
  if (i >= d.H) return;  // whole warp
  const uint32_t pxb = (uint32_t)d.C * 4u;  // bytes per pixel
  if (DO_GF) {
    Geo g;
    make_geo<true>(d, fx, fy, i, min(j, d.W - 1), g);
    s_off[warp][lane] = make_uint4((uint32_t)(g.y0 * d.W + g.x0) * pxb, (uint32_t)(g.y0 * d.W + g.x1) * pxb,
                                   (uint32_t)(g.y1 * d.W + g.x0) * pxb, (uint32_t)(g.y1 * d.W + g.x1) * pxb);
    const int ok = (int)g.oknw | ((int)g.okne << 1) | ((int)g.oksw << 2) | ((int)g.okse << 3) |
                   ((int)g.clipx << 4) | ((int)g.clipy << 5);
    s_aux[warp][lane] = make_float4(g.ax, g.ay, m, __int_as_float(ok));
  }
  __syncwarp();
  const int lq = lane % LP, grp = lane / LP;
  const int npx = min(TW, d.W - bx * TW);
  const int nq = QI > 0 ? QI : ((p.cchunk >> 2) - lq + LP - 1) / LP;
  const float det_scale_m = (DET && DO_GX) ? fixed_scale_from(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]),
                                                              p.count_log2) : 1.f;
  // (a power of two: exact; NaN when gout holds a non-finite value)
  const float det_inv_m = (DET && DO_GX) ? fixed_inv_scale(__uint_as_float(p.maxbits[0]) * __uint_as_float(p.maxbits[1]), det_scale_m) : 1.f;
  // blockIdx.y: channel slice of p.cchunk channels (small levels: more CTAs than tiles; every slice bins the
  // tile again, the slices' grad-flow / grad-mask partial sums are added by sum_parts_kernel)
  const uint32_t cb0 = blockIdx.y * (uint32_t)p.cchunk * 4u + lq * 16;
  const char* gl = reinterpret_cast<const char*>(p.gout) + cb0;  // + 16 * source key
  const char* xl = reinterpret_cast<const char*>(p.x) + (int64_t)(n % d.x_batch) * HW * pxb + cb0;
  const int64_t rowpix = (int64_t)n * HW + (int64_t)i * d.W + bx * TW + grp;  // this lane group's first pixel
  char* gxl = DO_GX ? reinterpret_cast<char*>(p.gx) + rowpix * pxb + cb0 : nullptr;
  const char* gol = reinterpret_cast<const char*>(p.gout) + rowpix * pxb + cb0;  // own gout (DO_GF: n is the frame)
#pragma unroll 1
  for (int s = 0; s < npx; s += G) {
    const int pa = s + grp;
    const bool act = (G == 1) || (pa < npx);
    constexpr int NQ = QI > 0 ? QI : 1;
    float sa = 0.f, sb = 0.f, sc = 0.f, se = 0.f;  // sum_c gout[c] * x_corner[c]
    if (QI > 0) {
      int cnt = 0;
      int4 e0 = make_int4(0, 0, 0, 0), e1 = e0;
      float4 v[4][NQ];
      DotRegs<NQ> dr;
      if (act) {
        if (DO_GX) {
          cnt = (DET && (s_cnt[warp][pa] & 0x40000000)) ? 0 : min(s_cnt[warp][pa] & 0xffffff, CAP);
          e0 = s_ent[0][warp][pa];
          e1 = s_ent[1][warp][pa];
          gx_issue<LP, NQ>(gl, cnt, e0, e1, v);
        }
        if (DO_GF) dot_issue_x<LP, NQ>(xl, s_off[warp][pa], dr);
        // (deterministic mode: hot destinations are left to the integer pass after this loop)
        if (DO_GX && !(DET && (s_cnt[warp][pa] & 0x40000000)))
          gx_finish<LP, NQ, NP>(gl, gxl, cnt, e0, e1, &s_ent[0][warp][pa], TH * TW, v, do_gx);
        if (DO_GF) {
          // the pixel's own gout row (an L1 / L2 hit: its neighbours just gathered it) is fetched last, which
          // keeps the merged batch inside the register budget of four CTAs per SM (64 registers)
          dot_issue_g<LP, NQ>(gol, dr);
          dot_finish<NQ>(dr, sa, sb, sc, se);
        }
      }
    } else if (act) {
#pragma unroll 1
      for (int qi = 0; qi < nq; ++qi) {
        if (DO_GX && !(DET && (s_cnt[warp][pa] & 0x40000000))) {
          const int cnt = min(s_cnt[warp][pa] & 0xffffff, CAP);
          const int4 e0 = s_ent[0][warp][pa], e1 = s_ent[1][warp][pa];
          float4 v[4][1];
          gx_issue<LP, 1>(gl + qi * (LP * 16), cnt, e0, e1, v);
          gx_finish<LP, 1, NP>(gl + qi * (LP * 16), gxl + qi * (LP * 16), cnt, e0, e1, &s_ent[0][warp][pa], TH * TW, v,
                               do_gx);
        }
        if (DO_GF) {
          DotRegs<1> dr;
          dot_issue<LP, 1>(xl + qi * (LP * 16), gol + qi * (LP * 16), s_off[warp][pa], dr);
          dot_finish<1>(dr, sa, sb, sc, se);
        }
      }
    }
    if (DO_GF) {
      // the four dot products of the pixel, summed over its LP lanes; the per-pixel algebra that turns them into
      // grad-flow / grad-mask runs once per row below (lane = pixel), not once per step on every lane
      const float r = reduce4<LP>(sa, sb, sc, se, lq);
      if (act && (lq % (LP / 4)) == 0) reinterpret_cast<float*>(DET ? reinterpret_cast<float4*>(&s_off[warp][pa]) : &s_sum[DET ? 0 : warp][pa])[lq / (LP / 4)] = r;
    }
    if (DO_GX) gxl += G * pxb;
    gol += G * pxb;
  }
  if (DO_GX && DET && do_gx) {
    // ---- deterministic mode, the hot destinations of this row: order-independent 64-bit fixed-point sums of the
    // list entries, plus the accumulator row when other mechanisms added terms there (a separate pass: they are rare,
    // and the float and the integer code never compete for registers)
    if (__ballot_sync(0xffffffffu, live && (s_cnt[warp][lane] & 0x40000000))) {
      char* gx0 = reinterpret_cast<char*>(p.gx) + rowpix * pxb + cb0;
#pragma unroll 1
      for (int s = 0; s < npx; s += G) {
        const int pa = s + grp;
        const int sc = ((G == 1) || (pa < npx)) ? s_cnt[warp][pa] : 0;
        if (sc & 0x40000000) {
          const int cnt = min(sc & 0xffffff, CAP);
          const int4 e0 = s_ent[0][warp][pa], e1 = s_ent[1][warp][pa];
#pragma unroll 1
          for (int qi = 0; qi < nq; ++qi) {
            float4 v[4][1];
            gx_issue<LP, 1>(gl + qi * (LP * 16), cnt, e0, e1, v);
            const long long* ha = p.acc64 + (rowpix + s) * (int64_t)d.C + (lq + qi * LP) * 4;
            gx_finish_det<LP, 1, NP>(gl + qi * (LP * 16), gx0 + (int64_t)s * pxb + qi * (LP * 16), cnt, e0, e1,
                                     &s_ent[0][warp][pa], TH * TW, v, det_scale_m, det_inv_m,
                                     (sc & 0x20000000) ? ha : nullptr);
          }
        }
      }
    }
  }
  if (DO_GF) {
    __syncwarp();
    if (live) {
      const float4 sum = DET ? *reinterpret_cast<const float4*>(&s_off[warp][lane]) : s_sum[DET ? 0 : warp][lane];
      const float4 aux = s_aux[warp][lane];  // ax, ay, mask, flags
      const int ok = __float_as_int(aux.w);
      // corners outside the image contribute nothing (ATen within_bounds)
      const float qa = (ok & 1) ? sum.x : 0.f, qb = (ok & 2) ? sum.y : 0.f;
      const float qc = (ok & 4) ? sum.z : 0.f, qe = (ok & 8) ? sum.w : 0.f;
      // bilinear weights from the fractions: x1 - ix == 1 - ax and ix - x0 == ax, the expressions of make_geo
      const float bxw = 1.f - aux.x, byw = 1.f - aux.y;
      const float gix = (qb - qa) * byw + (qe - qc) * aux.y;
      const float giy = (qc - qa) * bxw + (qe - qb) * aux.x;
      const float gm = fmaf(qe, aux.x * aux.y, fmaf(qc, bxw * aux.y, fmaf(qb, aux.x * byw, qa * (bxw * byw))));
      const float mm = HAS_MASK ? aux.z : 1.f;  // the sums used gout, not gout*mask
      // d(clipped coordinate)/d(flow), as make_geo forms it: clip-grad * size/2 * 1/((size-1)/2)
      const float gmx = ((ok & 16) ? 0.f : 1.f) * (0.5f * (float)d.W) * d.inv_bw;
      const float gmy = ((ok & 32) ? 0.f : 1.f) * (0.5f * (float)d.H) * d.inv_bh;
      float* gfp = p.gflow;
      float* gmp = p.gmask;
      if (gridDim.y > 1) {
        gfp = p.gpart + (int64_t)blockIdx.y * 3 * d.N * HW;
        gmp = gfp + (int64_t)2 * d.N * HW;
      }
      if (gfp) {
        float* gf = gfp + (int64_t)n * 2 * HW + pix;
        gf[0] = gix * mm * gmx;
        gf[HW] = giy * mm * gmy;
      }
      if (gmp) gmp[(int64_t)n * HW + pix] = gm;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// NCHW gather: one thread per pixel, channel loop.
template <int TH, int TW, bool DO_GX, bool DO_GF, bool HAS_MASK, bool USE_TMA, bool REPEAT>
__global__ void __launch_bounds__(TH* TW, 2) gather_nchw_kernel(const __grid_constant__ BwdParams p,
                                                                const __grid_constant__ CUtensorMap tm_flow,
                                                                const __grid_constant__ CUtensorMap tm_mask) {
  __shared__ TileSmem<TH, TW> s;
  const Dims& d = p.d;
  const int tid = threadIdx.x;
  const int tx = tid % TW, ty = tid / TW;
  const int tiles_x = (d.W + TW - 1) / TW, tiles_y = (d.H + TH - 1) / TH;
  const int nimg = DO_GF ? d.N : d.x_batch;
  const int total = nimg * tiles_y * tiles_x;
  const int HW = d.H * d.W;
  const int c0 = blockIdx.y * p.cchunk;
  const int nc = min(p.cchunk, d.C - c0);

  if (DO_GF && USE_TMA)
    tile_pipeline_init<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, blockIdx.x, total, tiles_x, tiles_y);
  int buf = 0;
  uint32_t phases = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int bx = t % tiles_x;
    const int r = t / tiles_x;
    const int by = r % tiles_y;
    const int n = r / tiles_y;
    const int i = by * TH + ty, j = bx * TW + tx;
    const bool live = (i < d.H) & (j < d.W);
    const int pix = i * d.W + j;
    float fx = 0.f, fy = 0.f, m = 1.f;
    if (DO_GF) {
      if (USE_TMA) {
        const int tn = t + gridDim.x;
        if (tid == 0 && tn < total) issue_tile<TH, TW, HAS_MASK>(s, &tm_flow, &tm_mask, tn, tiles_x, tiles_y, buf ^ 1);
        mbar_wait(&s.bar[buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
        fx = s.flow[buf][0][ty][tx];
        fy = s.flow[buf][1][ty][tx];
        if (HAS_MASK) m = s.mask[buf][ty][tx];
      } else if (live) {
        const float* fl = p.flow + (int64_t)n * 2 * HW + pix;
        fx = __ldg(fl);
        fy = __ldg(fl + HW);
        if (HAS_MASK) m = __ldg(p.mask + (int64_t)n * HW + pix);
      }
    }
    if (live) {
      // ---- destination role: contributor list into registers
      int cnt = 0;
      int64_t soff[kListCap];
      float sw[kListCap];
      if (DO_GX) {
        const int64_t D = (int64_t)n * HW + pix;
        const int64_t ndest = (int64_t)HW * d.x_batch;
        cnt = min(__ldg(p.cnt + D), kListCap);
        const int4* ep = reinterpret_cast<const int4*>(p.entries) + D;
#pragma unroll
        for (int k = 0; k < kListCap; k += 2) {
          if (k < cnt) {
            const int4 raw = __ldg(ep + (int64_t)(k >> 1) * ndest);
            const int s0 = raw.x, s1 = raw.z;
            if (REPEAT) {
              soff[k] = (int64_t)(s0 / HW) * d.C * HW + (s0 % HW);
              soff[k + 1] = (int64_t)(s1 / HW) * d.C * HW + (s1 % HW);
            } else {
              soff[k] = s0 - n * HW;  // same frame: pixel offset inside the frame
              soff[k + 1] = s1 - n * HW;
            }
            sw[k] = __int_as_float(raw.y);
            sw[k + 1] = __int_as_float(raw.w);
          }
        }
      }
      Geo g = {};
      if (DO_GF) make_geo<true>(d, fx, fy, i, j, g);
      const float* gfr = p.gout + (REPEAT ? (int64_t)0 : (int64_t)n * d.C * HW) + (int64_t)c0 * HW;  // list base
      const float* gop = p.gout + ((int64_t)n * d.C + c0) * HW + pix;
      float* gxp = DO_GX ? p.gx + ((int64_t)n * d.C + c0) * HW + pix : nullptr;
      const float* xc = p.x + ((int64_t)(n % d.x_batch) * d.C + c0) * HW;
      const float* pnw = xc + (g.y0 * d.W + g.x0);
      const float* pne = xc + (g.y0 * d.W + g.x1);
      const float* psw = xc + (g.y1 * d.W + g.x0);
      const float* pse = xc + (g.y1 * d.W + g.x1);
      float gix = 0.f, giy = 0.f, gm = 0.f;
      if (!DO_GX && DO_GF && nc == 3) {
        // image-like tensors (the C = 3 warps of the loss / preview / kitti sites, where only the flow carries a
        // gradient): the fifteen values of the pixel are requested back to back, then used in channel order (the
        // same sums, bit for bit, as the loop below -- one exposed memory round trip instead of two)
        float go[3], v[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          go[c] = __ldg(gop + (int64_t)c * HW);
          v[c][0] = __ldg(pnw + (int64_t)c * HW);
          v[c][1] = __ldg(pne + (int64_t)c * HW);
          v[c][2] = __ldg(psw + (int64_t)c * HW);
          v[c][3] = __ldg(pse + (int64_t)c * HW);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float vnw = g.oknw ? v[c][0] : 0.f, vne = g.okne ? v[c][1] : 0.f;
          const float vsw = g.oksw ? v[c][2] : 0.f, vse = g.okse ? v[c][3] : 0.f;
          gix = fmaf(go[c], (vne - vnw) * (1.f - g.ay) + (vse - vsw) * g.ay, gix);
          giy = fmaf(go[c], (vsw - vnw) * (1.f - g.ax) + (vse - vne) * g.ax, giy);
          gm = fmaf(go[c], fmaf(vse, g.wse, fmaf(vsw, g.wsw, fmaf(vne, g.wne, vnw * g.wnw))), gm);
        }
      } else
#pragma unroll 2
      for (int c = 0; c < nc; ++c) {
        if (DO_GX) {
          float v[kListCap];
#pragma unroll
          for (int k = 0; k < kListCap; ++k)
            if (k < cnt) v[k] = __ldg(gfr + soff[k]);
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < kListCap; ++k)
            if (k < cnt) acc = fmaf(sw[k], v[k], acc);
          st_stream(gxp, acc);
          gxp += HW;
          gfr += HW;
        }
        if (DO_GF) {
          const float go = __ldg(gop);
          float vnw = __ldg(pnw), vne = __ldg(pne), vsw = __ldg(psw), vse = __ldg(pse);
          vnw = g.oknw ? vnw : 0.f;
          vne = g.okne ? vne : 0.f;
          vsw = g.oksw ? vsw : 0.f;
          vse = g.okse ? vse : 0.f;
          gix = fmaf(go, (vne - vnw) * (1.f - g.ay) + (vse - vsw) * g.ay, gix);
          giy = fmaf(go, (vsw - vnw) * (1.f - g.ax) + (vse - vne) * g.ax, giy);
          gm = fmaf(go, fmaf(vse, g.wse, fmaf(vsw, g.wsw, fmaf(vne, g.wne, vnw * g.wnw))), gm);
          gop += HW;
          pnw += HW;
          pne += HW;
          psw += HW;
          pse += HW;
        }
      }
      if (DO_GF) {
        // channel chunks (blockIdx.y) are only used by the launcher when gflow/gmask are not
        // requested, so these are complete sums
        const float mm = HAS_MASK ? m : 1.f;
        if (p.gflow) {
          float* gf = p.gflow + (int64_t)n * 2 * HW + pix;
          gf[0] = gix * mm * g.gmx;
          gf[HW] = giy * mm * g.gmy;
        }
        if (p.gmask) p.gmask[(int64_t)n * HW + pix] = gm;
      }
    }
    if (DO_GF && USE_TMA) {
      __syncthreads();
      buf ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct GatherWs {
  int* cnt;        // [x_batch*H*W] + the overflow-list length right behind it (one memset clears both)
  int* ovf_count;
  void* entries;
  unsigned char* ovf;
  int* ovf_list;
  size_t cnt_bytes;
  size_t bytes;
};

static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
int memset_failed();  // warp_bwd.cu: sets the error text

// global-list scheme (NCHW kernels)
static GatherWs carve(void* base, int64_t N, int H, int W, int64_t x_batch) {
  const size_t npix_d = (size_t)x_batch * H * W, npix_o = (size_t)N * H * W;
  GatherWs w;
  char* b = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.cnt = reinterpret_cast<int*>(b + o);
  w.ovf_count = w.cnt + npix_d;
  w.cnt_bytes = (npix_d + 1) * sizeof(int);
  o += up256(w.cnt_bytes);
  w.entries = b + o;
  o += up256(npix_d * kListCap * sizeof(ListEntry));
  w.ovf = reinterpret_cast<unsigned char*>(b + o);
  o += up256(npix_o);
  w.ovf_list = reinterpret_cast<int*>(b + o);
  o += up256(npix_o * sizeof(int));
  w.bytes = o;
  return w;
}

// local-binning scheme (channels-last kernels): [tcnt | ovf_count | ovf flags] are cleared by one memset
struct LocalWs {
  int* tcnt;
  int* ovf_count;
  int *bcount, *bfill, *iseg_count, *flex_count, *flex_next;  // cleared with tcnt
  int *bstart, *iseg_list, *flex_list, *pool;
  unsigned char* ovf;
  int2* tlist;
  int* ovf_list;
  int4* pixrec;
  float* gpart;
  int slices;
  size_t ovf_stride;
  int cand_cap;
  size_t clear_bytes;
  size_t bytes;
  // deterministic mode only
  unsigned* maxbits;
  unsigned char* touched;
  long long* acc64;
  int* dcnt;               // in-image corners per destination (bit 30: a contribution was handed to overflow_kernel)
  int* incoh;              // incoherent segments per image
  size_t det_clear_bytes;  // [maxbits | touched | dcnt | incoh]
};

static LocalWs carve_local(void* base, int64_t N, int H, int W, int64_t x_batch, int C, bool det = false) {
  const size_t ntile = (size_t)x_batch * ((H + 7) / 8) * ((W + 31) / 32), npix_o = (size_t)N * H * W;
  // small level (fewer tiles than CTA slots): room for a channel-sliced gather -- per-slice overflow flags
  // and list entries, per-slice grad-flow / grad-mask partial sums
  const int slices = det ? 1 : channel_slices(N, C, H, W);
  const bool small = slices > 1;
  const size_t nov = small ? 1 + slices : 1;
  LocalWs w;
  int64_t cap = (int64_t)kCandPerFrame * (x_batch > 0 ? N / x_batch : 1);
  w.cand_cap = (int)(cap > kCandMax ? kCandMax : cap);
  char* b = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.tcnt = reinterpret_cast<int*>(b + o);
  w.ovf_count = w.tcnt + ntile;
  w.bcount = w.ovf_count + 1;
  w.bfill = w.bcount + ntile;
  w.iseg_count = w.bfill + ntile;
  w.flex_count = w.iseg_count + 1;
  w.flex_next = w.flex_count + 1;
  o += up256((3 * ntile + 4) * sizeof(int));
  w.ovf = reinterpret_cast<unsigned char*>(b + o);
  w.ovf_stride = up256(npix_o);
  o += nov * w.ovf_stride;
  w.clear_bytes = o;
  w.tlist = reinterpret_cast<int2*>(b + o);
  o += up256(ntile * w.cand_cap * sizeof(int2));
  w.ovf_list = reinterpret_cast<int*>(b + o);
  o += up256(nov * npix_o * sizeof(int));
  w.pixrec = reinterpret_cast<int4*>(b + o);
  o += up256(npix_o * sizeof(int4));
  w.bstart = w.flex_list = w.iseg_list = w.pool = nullptr;
  if (det) {  // the counting sort of incoherent segments serves the deterministic mode only
    w.bstart = reinterpret_cast<int*>(b + o);
    o += up256(ntile * sizeof(int));
    w.flex_list = reinterpret_cast<int*>(b + o);
    o += up256(ntile * sizeof(int));
    w.iseg_list = reinterpret_cast<int*>(b + o);
    o += up256((size_t)N * H * ((W + 31) / 32) * sizeof(int));
    w.pool = reinterpret_cast<int*>(b + o);
    o += up256(4 * npix_o * sizeof(int));
  }
  w.gpart = nullptr;
  w.slices = slices;
  if (small) {
    w.gpart = reinterpret_cast<float*>(b + o);
    o += up256((size_t)slices * 3 * npix_o * sizeof(float));
  }
  w.maxbits = nullptr;
  w.touched = nullptr;
  w.acc64 = nullptr;
  w.dcnt = nullptr;
  w.incoh = nullptr;
  w.det_clear_bytes = 0;
  if (det) {
    const size_t npix_d = (size_t)x_batch * H * W, o0 = o;
    w.maxbits = reinterpret_cast<unsigned*>(b + o);
    o += 256;
    w.touched = reinterpret_cast<unsigned char*>(b + o);
    o += up256(npix_d);
    w.dcnt = reinterpret_cast<int*>(b + o);
    o += up256(npix_d * sizeof(int));
    w.incoh = reinterpret_cast<int*>(b + o);
    o += up256((size_t)x_batch * sizeof(int));
    w.det_clear_bytes = o - o0;
    // not cleared as a whole: zero_hot_rows_kernel clears the rows that can receive terms
    w.acc64 = reinterpret_cast<long long*>(b + o);
    o += up256(npix_d * (size_t)C * sizeof(long long));
  }
  w.bytes = o;
  return w;
}

size_t local_det_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch) {
  return carve_local(nullptr, N, H, W, x_batch, C, true).bytes;
}

// the caller does not tell the layout when it asks: size for the larger (global-list) scheme
size_t gather_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch) {
  const size_t a = carve(nullptr, N, H, W, x_batch).bytes, b = carve_local(nullptr, N, H, W, x_batch, C).bytes;
  return a > b ? a : b;
}

bool gather_supported(const BwdParams& p, Layout lx, Layout lg) {
  const Dims& d = p.d;
  if (d.flags & (C2M_FLAG_BWD_ATOMIC | C2M_FLAG_FORCE_GENERIC | C2M_FLAG_COORD_GRID | C2M_FLAG_TRUE_DIV | C2M_FLAG_NO_FMA))
    return false;
  // deterministic grad-input: only the channels-last gather has an order-independent form
  if ((d.flags & C2M_FLAG_DETERMINISTIC) && p.gx && lx != LAYOUT_NHWC) return false;
  if (p.other || p.gother) return false;
  if (lx != lg || lx == LAYOUT_OTHER) return false;
  // NCHW image-like tensors (C = 3): building contributor lists costs more than the few atomics they save
  // (measured at 40 x 3 x 256 x 512: 0.24 ms with the direct scatter, 0.63 ms with lists)
  if (lx == LAYOUT_NCHW && p.gx && d.C < 8) return false;
  if ((int64_t)d.N * d.H * d.W >= (1ll << 31) - 1) return false;
  if (lx == LAYOUT_NHWC) {
    if ((d.C & 3) || ((uintptr_t)p.x & 15) || ((uintptr_t)p.gout & 15) || (p.gx && ((uintptr_t)p.gx & 15)))
      return false;
    if ((int64_t)d.H * d.W * d.C >= (1ll << 30)) return false;  // 32-bit byte offsets inside one image
    if (d.H >= 32768 || d.W >= 32768) return false;              // 16-bit corner coordinates in the pixel records
    if (d.N > 65535) return false;                                // segbin_kernel puts the frame in gridDim.y
    if ((int64_t)d.N * d.H * d.W * (d.C / 4) >= (1ll << 32)) return false;  // 32-bit source keys (16-byte units)
  }
  return true;
}

// channel-sliced gather: grad-flow / grad-mask = sum of the slices' partial sums, in slice order
__global__ void __launch_bounds__(256) sum_parts_kernel(const float* __restrict__ part, int slices, int64_t total,
                                                        int64_t nflow, float* __restrict__ gflow,
                                                        float* __restrict__ gmask) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  float s = part[k];
  for (int c = 1; c < slices; ++c) s += part[c * total + k];
  if (k < nflow) {
    if (gflow) gflow[k] = s;
  } else if (gmask) {
    gmask[k - nflow] = s;
  }
}

// C2M_WARP_FLEX=0: incoherent segments take the overflow list (atomics) instead of the per-tile counting sort
static bool flex_enabled() {
  static const bool v = [] {
    const char* e = getenv("C2M_WARP_FLEX");
    return !(e && *e == '0');
  }();
  return v;
}

template <int LP, int QI, bool DO_GX, bool DO_GF, bool DET>
static void launch_gather_nhwc(BwdParams p, cudaStream_t st) {
  p.pf_tiles = prefetch_tiles(-1);  // measured: with four CTAs per SM resident the L2 prefetch gains nothing
  constexpr int TH = 8, TW = 32;
  const Dims& d = p.d;
  const int tiles = p.nframes * ((d.H + TH - 1) / TH) * ((d.W + TW - 1) / TW);
  TileMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (DO_GF) tm = make_tile_maps(d, p.flow, p.mask, TH, TW);
  const int slices = d.C / p.cchunk;
#define C2M_LAUNCH(MASK, TMA)                                                                                    \
  gather_nhwc_kernel<LP, QI, DO_GX, DO_GF, MASK, TMA, DET><<<dim3(tiles, slices), TH * TW, 0, st>>>(p, tm.flow, \
                                                                                                     tm.mask)
  if (p.mask) {
    if (tm.ok) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tm.ok) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
  if (DO_GF && slices > 1) {
    const int64_t nflow = (int64_t)2 * d.N * d.H * d.W, total = nflow + nflow / 2;
    sum_parts_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.gpart, slices, total, nflow, p.gflow, p.gmask);
    count_launch();
  }
}

template <bool DO_GX, bool DO_GF, bool DET = false>
static void launch_gather_nhwc_lp(BwdParams p, cudaStream_t st) {
  const int C4 = p.cchunk / 4;  // channels of one blockIdx.y slice (launch_bwd_gather decides; normally all)
  switch (C4) {  // two float4 groups per lane where C allows: half the per-pixel overhead of one
    case 8: return launch_gather_nhwc<4, 2, DO_GX, DO_GF, DET>(p, st);    // C = 32
    case 16: return launch_gather_nhwc<8, 2, DO_GX, DO_GF, DET>(p, st);   // C = 64
    case 32: return launch_gather_nhwc<16, 2, DO_GX, DO_GF, DET>(p, st);  // C = 128
    case 64: return launch_gather_nhwc<32, 2, DO_GX, DO_GF, DET>(p, st);  // C = 256
    default: break;
  }
  // any other channel count: run-time channel loop
  if (C4 >= 24) return launch_gather_nhwc<32, 0, DO_GX, DO_GF, DET>(p, st);
  if (C4 >= 12) return launch_gather_nhwc<16, 0, DO_GX, DO_GF, DET>(p, st);
  if (C4 >= 6) return launch_gather_nhwc<8, 0, DO_GX, DO_GF, DET>(p, st);
  return launch_gather_nhwc<4, 0, DO_GX, DO_GF, DET>(p, st);
}

template <bool DO_GX, bool DO_GF, bool REPEAT>
static void launch_gather_nchw(BwdParams p, cudaStream_t st) {
  constexpr int TH = 8, TW = 32;
  const Dims& d = p.d;
  const int nimg = DO_GF ? d.N : d.x_batch;
  const int tiles = nimg * ((d.H + TH - 1) / TH) * ((d.W + TW - 1) / TW);
  int ysplit = 1;
  if (!DO_GF) {  // channel chunks only when no per-pixel reduction over channels is produced
    const int want = sm_count() * 4;
    while (tiles * ysplit < want && (d.C / (ysplit * 2)) >= 8) ysplit *= 2;
  }
  p.cchunk = (d.C + ysplit - 1) / ysplit;
  ysplit = (d.C + p.cchunk - 1) / p.cchunk;
  TileMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (DO_GF) tm = make_tile_maps(d, p.flow, p.mask, TH, TW);
#define C2M_LAUNCH(MASK, TMA)                                                                              \
  do {                                                                                                        \
    auto kfn = gather_nchw_kernel<TH, TW, DO_GX, DO_GF, MASK, TMA, REPEAT>;                                \
    int cap = resident_ctas(reinterpret_cast<const void*>(kfn), TH * TW) / ysplit;                         \
    if (cap < 1) cap = 1;                                                                                  \
    kfn<<<dim3(tiles < cap ? tiles : cap, ysplit), TH * TW, 0, st>>>(p, tm.flow, tm.mask);                 \
  } while (0)
  if (p.mask) {
    if (tm.ok) C2M_LAUNCH(true, true); else C2M_LAUNCH(true, false);
  } else {
    if (tm.ok) C2M_LAUNCH(false, true); else C2M_LAUNCH(false, false);
  }
#undef C2M_LAUNCH
  count_launch();
}

// max |a[i]| over a dense array, as the bit pattern of a non-negative float (order independent)
__global__ void __launch_bounds__(256) absmax_flat_kernel(const float* __restrict__ a, int64_t n, unsigned* out_bits) {
  float mx = 0.f;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(a) & 15) ? 0 : n >> 2;  // vector loads only when aligned
  const float4* a4 = reinterpret_cast<const float4*>(a);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(a4 + k);
    // (fmaxf drops a NaN operand: test the sum, which is NaN as soon as one element is)
    const float m4 = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    const float chk = v.x + v.y + v.z + v.w;
    mx = (chk == chk) ? fmaxf(mx, m4) : __int_as_float(0x7f800000);
  }
  for (int64_t k = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(a[k]);
    mx = (v == v) ? fmaxf(mx, v) : __int_as_float(0x7f800000);  // NaN counts as +inf: "non-finite seen"
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out_bits, __float_as_uint(mx));
}

__global__ void set_bits_kernel(unsigned* p, unsigned v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// Plan (c2m_warp_plan): the segment registration of the channels-last float backward depends on the flow and the mask
// only, so the forward call can run it -- on a second stream, next to the HBM-bound forward kernel (segbin_kernel is
// bound by integer instructions: 0.07 ms of the headline backward that then cost nothing) -- into the buffer the
// backward later receives as its workspace with C2M_FLAG_PLANNED.
static void bind_local(BwdParams& p, const LocalWs& w) {
  p.tcnt = w.tcnt;
  p.bcount = w.bcount; p.bstart = w.bstart; p.bfill = w.bfill; p.pool = w.pool;
  p.iseg_count = w.iseg_count; p.iseg_list = w.iseg_list;
  p.flex_count = w.flex_count; p.flex_list = w.flex_list; p.flex_next = w.flex_next;
  p.tlist = w.tlist;
  p.cand_cap = w.cand_cap;
  p.pixrec = w.pixrec;
  p.gpart = w.gpart;
  p.ovf_stride = (int64_t)w.ovf_stride;
  p.cchunk = p.d.C / w.slices;  // small pyramid level: every pass of this call uses the same channel slicing
  p.ovf = w.ovf;
  p.ovf_count = w.ovf_count;
  p.ovf_list = w.ovf_list;
}

static bool plan_supported(const Dims& d) {
  if (d.flags & (C2M_FLAG_BWD_ATOMIC | C2M_FLAG_FORCE_GENERIC | C2M_FLAG_COORD_GRID | C2M_FLAG_TRUE_DIV | C2M_FLAG_NO_FMA |
                 C2M_FLAG_DETERMINISTIC | C2M_FLAG_STAGE_NHWC))
    return false;
  if (d.N <= 0 || d.C <= 0 || d.H <= 0 || d.W <= 0) return false;
  // the channels-last conditions of gather_supported() that do not involve the tensors' addresses
  if ((int64_t)d.N * d.H * d.W >= (1ll << 31) - 1 || (d.C & 3)) return false;
  if ((int64_t)d.H * d.W * d.C >= (1ll << 30) || d.H >= 32768 || d.W >= 32768 || d.N > 65535) return false;
  if ((int64_t)d.N * d.H * d.W * (d.C / 4) >= (1ll << 32)) return false;
  return true;
}

size_t plan_bytes(const Dims& d) {
  return plan_supported(d) ? 256 + carve_local(nullptr, d.N, d.H, d.W, d.x_batch, d.C, false).bytes : 0;
}

int launch_plan(const BwdParams& pin, void* plan, size_t bytes, cudaStream_t st) {
  BwdParams p = pin;
  const Dims& d = p.d;
  const size_t need = plan_bytes(d);
  if (need == 0) {
    set_error("c2m_warp_plan: this configuration has no plan (c2m_warp_plan_bytes() == 0)");
    return C2M_ERR_INVALID;
  }
  if (!plan || bytes < need) {
    set_error("plan buffer too small: %zu < %zu", bytes, need);
    return C2M_ERR_WORKSPACE;
  }
  const LocalWs w = carve_local(reinterpret_cast<char*>(plan) + 256, d.N, d.H, d.W, d.x_batch, d.C, false);
  p.n0 = 0;
  p.nframes = d.N;
  p.key_mul = d.C / 4;
  bind_local(p, w);
  p.bcount = nullptr;  // float path: incoherent segments take the overflow list
  p.cnt = nullptr;
  if (cudaMemsetAsync(w.tcnt, 0, w.clear_bytes, st) != cudaSuccess) return memset_failed();
  const int segs = d.H * ((d.W + 31) / 32);
  // 128-thread blocks: next to a running forward kernel (which leaves 4 K registers and 512 threads of an SM free) one
  // still fits.  (A persistent grid of 1-4 such blocks per SM launched before the forward kernel was measured too: the
  // registration then takes 0.65-0.98 ms and becomes the critical path of the forward call.)
  constexpr int wpb = 4;
  segbin_kernel<false><<<dim3((unsigned)((segs + wpb - 1) / wpb), (unsigned)d.N), wpb * 32, 0, st>>>(p);
  count_launch();
  return C2M_OK;
}

int launch_bwd_gather(const BwdParams& pin, Layout lx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  BwdParams p = pin;
  const Dims& d = p.d;
  const bool need_gf = p.gflow || p.gmask;
  const bool repeat = d.x_batch != d.N;
  const bool fuse = p.gx && need_gf && !repeat;
  const bool local = lx == LAYOUT_NHWC;  // channels-last: local binning; NCHW kernels: global contributor lists
  p.n0 = 0;
  p.nframes = d.N;
  p.key_mul = lx == LAYOUT_NHWC ? d.C / 4 : 1;  // channels-last lists address gout in 16-byte units
  p.cchunk = d.C;
  p.gpart = nullptr;
  p.ovf_stride = 0;
  const bool det = (d.flags & C2M_FLAG_DETERMINISTIC) && p.gx;
  if (p.gx && local) {
    const LocalWs w = carve_local(workspace, d.N, d.H, d.W, d.x_batch, d.C, det);
    if (!workspace || workspace_bytes < w.bytes) {
      set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes);
      return C2M_ERR_WORKSPACE;
    }
    bind_local(p, w);
    // C2M_FLAG_PLANNED: the forward call has already cleared the counters and run segbin_kernel on this workspace
    const bool planned = (d.flags & C2M_FLAG_PLANNED) && !det;
    if (!planned && cudaMemsetAsync(w.tcnt, 0, w.clear_bytes, st) != cudaSuccess) return memset_failed();
    if (det) {
      // fixed-point scale from max|gout| * max|mask| (a bound of every |term|), accumulator rows cleared
      p.maxbits = w.maxbits;
      p.touched = w.touched;
      p.acc64 = w.acc64;
      p.cnt = w.dcnt;
      p.incoh = w.incoh;
      // an image whose incoherent segments exceed 1/16 of all its segments gets every accumulator row cleared
      p.incoh_thresh = (int)(((int64_t)d.H * ((d.W + 31) / 32) * (d.N / d.x_batch)) / 16);
      int cl = 2;  // 4 corners
      const int64_t cnt = (int64_t)d.H * d.W * (d.N / d.x_batch);
      while ((1ll << (cl - 2)) < cnt) ++cl;
      p.count_log2 = cl;
      if (cudaMemsetAsync(w.maxbits, 0, w.det_clear_bytes, st) != cudaSuccess) return memset_failed();
      // the two maxima are formed by segbin_kernel on its way (it reads the mask anyway and streams gout next to
      // its integer work); without a mask the second factor is 1
      p.maxacc = w.maxbits;
      if (!p.mask) {
        set_bits_kernel<<<1, 1, 0, st>>>(w.maxbits + 1, 0x3f800000u);
        count_launch();
      }
    }
    // the counting sort of incoherent segments serves the deterministic mode (it replaces 64-bit atomics per
    // contribution); the float path keeps overflow_kernel's vector reductions, which are faster than the sort
    if (!det || !flex_enabled()) p.bcount = nullptr;
    const int segs = d.H * ((d.W + 31) / 32);  // per frame; grid.y = frames (N <= 65535 checked by gather_supported)
    if (!planned) {
      if (p.maxacc) segbin_kernel<true><<<dim3((unsigned)((segs + 7) / 8), (unsigned)d.N), 256, 0, st>>>(p);
      else segbin_kernel<false><<<dim3((unsigned)((segs + 7) / 8), (unsigned)d.N), 256, 0, st>>>(p);
      count_launch();
    }
    if (p.bcount) {
      // incoherent flows: finish the counting sort of their pixels by destination tile (both kernels leave at once
      // when segbin_kernel met no incoherent segment)
      const int ntile = d.x_batch * ((d.H + 7) / 8) * ((d.W + 31) / 32);
      flex_scan_kernel<<<1, 1024, 0, st>>>(p, ntile);
      flex_fill_kernel<<<sm_count() * 8, 256, 0, st>>>(p);
      count_launch(2);
    }
    if (det) {  // incoherent segments / failed registrations first: the gather folds their rows in
      const int64_t ndest = (int64_t)d.x_batch * d.H * d.W;
      zero_hot_rows_kernel<<<(unsigned)((ndest + 255) / 256), 256, 0, st>>>(w.dcnt, w.incoh, p.incoh_thresh, w.acc64,
                                                                              ndest, d.H * d.W, d.C);
      overflow_kernel<true, true><<<sm_count() * 8, 256, 0, st>>>(p);
      count_launch(2);
    }
  } else if (p.gx) {
    const GatherWs w = carve(workspace, d.N, d.H, d.W, d.x_batch);
    if (!workspace || workspace_bytes < w.bytes) {
      set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes);
      return C2M_ERR_WORKSPACE;
    }
    p.cnt = w.cnt;
    p.entries = w.entries;
    p.ovf = w.ovf;
    p.ovf_count = w.ovf_count;
    p.ovf_list = w.ovf_list;
    if (cudaMemsetAsync(p.cnt, 0, w.cnt_bytes, st) != cudaSuccess) return memset_failed();
    const int64_t total = (int64_t)d.N * d.H * d.W;
    bin_kernel<<<(unsigned)((total + kBinPixelsPerBlock - 1) / kBinPixelsPerBlock), 256, 0, st>>>(p);
    count_launch();
  }
  profile_begin(st);  // the gather kernel(s): the dominant part of the backward
  if (lx == LAYOUT_NHWC) {
    BwdParams q = p;
    q.nframes = d.x_batch;  // a grad-input-only pass walks the images of x
    if (det) {
      if (fuse) {
        launch_gather_nhwc_lp<true, true, true>(p, st);
      } else {
        launch_gather_nhwc_lp<true, false, true>(q, st);
        if (need_gf) launch_gather_nhwc_lp<false, true>(p, st);
      }
    } else if (fuse) {
      launch_gather_nhwc_lp<true, true>(p, st);
    } else {
      if (p.gx) launch_gather_nhwc_lp<true, false>(q, st);
      if (need_gf) launch_gather_nhwc_lp<false, true>(p, st);
    }
  } else if (fuse) {
    launch_gather_nchw<true, true, false>(p, st);
  } else {
    if (p.gx) {
      if (repeat) launch_gather_nchw<true, false, true>(p, st);
      else launch_gather_nchw<true, false, false>(p, st);
    }
    if (need_gf) launch_gather_nchw<false, true, false>(p, st);
  }
  profile_end(st);
  if (p.gx && local && p.bcount) {
    const int C4 = d.C / 4, grid = sm_count() * 4;
#define C2M_FLEX(LP) gather_flex_kernel<LP, true><<<grid, 256, 0, st>>>(p)  /* (bcount is only set in deterministic mode) */
    if (C4 >= 32) C2M_FLEX(32);
    else if (C4 >= 16) C2M_FLEX(16);
    else if (C4 >= 8) C2M_FLEX(8);
    else C2M_FLEX(4);
#undef C2M_FLEX
    count_launch();
  }
  if (p.gx && !(local && det)) {
    const int grid = sm_count() * 8;
    if (lx == LAYOUT_NHWC)  // gather_supported() has checked C % 4 and the 16-byte alignment
      overflow_kernel<true><<<grid, 256, 0, st>>>(p);
    else
      overflow_kernel<false><<<grid, 256, 0, st>>>(p);
    count_launch();
  }
  return C2M_OK;
}

}  // namespace c2m
