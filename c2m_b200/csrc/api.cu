// api.cu -- the C ABI of libc2m_warp.so (see include/c2m_warp.h) and host-side plumbing:
// argument validation, layout classification, TMA descriptor encoding, error reporting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace c2m {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct ProfileState {
  bool on = false, valid = false;
  cudaEvent_t a = nullptr, b = nullptr;
};
// process-wide on purpose: autograd runs the backward call on its own thread, the reader is the main thread
static ProfileState g_prof;
static std::mutex g_prof_mu;

void profile_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.on) return;
  g_prof.valid = cudaEventRecord(g_prof.a, st) == cudaSuccess;
}
void profile_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.on || !g_prof.valid) return;
  g_prof.valid = cudaEventRecord(g_prof.b, st) == cudaSuccess;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
    cached_dev = dev;
  }
  return cached;
}

// Channels-last kernels slice the channels over blockIdx.y while a level has fewer tiles than this
// (C2M_WARP_SPLIT_TILES overrides; 0 switches the slicing off).
int split_tiles() {
  static const int v = [] {
    const char* e = getenv("C2M_WARP_SPLIT_TILES");
    return e && *e ? atoi(e) : kSplitTiles;
  }();
  return v;
}

int channel_slices(int64_t N, int C, int H, int W) {
  if ((C & 3) || N * H * W >= (1ll << 24)) return 1;  // (slice tags share a 32-bit word with the pixel index)
  const int64_t tiles = N * ((H + 7) / 8) * ((W + 31) / 32);
  int C4 = C / 4, s = 1;
  while (tiles * s < split_tiles() && s < kSplitMax && C4 % 2 == 0 && C4 / 2 >= 16) {
    C4 /= 2;
    s *= 2;
  }
  return s;
}

int prefetch_tiles(int dflt) {
  static const int v = [] {
    const char* e = getenv("C2M_WARP_PREFETCH_TILES");
    return e && *e ? atoi(e) : INT_MIN;
  }();
  return v == INT_MIN ? dflt : v;
}

int resident_ctas(const void* kernel, int threads) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * sm_count();
}

// cuTensorMapEncodeTiled is a driver-API symbol; it is resolved at run time so that the library
// has no link-time dependency on libcuda (it must load on a machine without a GPU driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

bool make_tensor_map_3d(CUtensorMap* tm, const float* base, int W, int H, int64_t planes, int box_w, int box_h,
                        int box_p) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (W & 3) != 0) return false;  // 16-byte base and pitches
  if (box_w > 256 || box_h > 256 || box_p > 256) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_p};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static Layout classify(const int64_t s[4], int64_t N, int C, int H, int W) {
  const int64_t HW = (int64_t)H * W;
  auto ok = [](int64_t size, int64_t stride, int64_t want) { return size == 1 || stride == want; };
  if (ok(C, s[1], HW) && ok(H, s[2], W) && ok(W, s[3], 1) && ok(N, s[0], (int64_t)C * HW)) return LAYOUT_NCHW;
  if (ok(C, s[1], 1) && ok(H, s[2], (int64_t)W * C) && ok(W, s[3], C) && ok(N, s[0], (int64_t)C * HW)) return LAYOUT_NHWC;
  return LAYOUT_OTHER;
}

static void canonical(int64_t dst[4], const int64_t src[4], Layout l, int C, int H, int W) {
  // size-1 dimensions may carry arbitrary strides; rewrite them so kernels can trust the numbers
  if (l == LAYOUT_NCHW) {
    dst[0] = (int64_t)C * H * W; dst[1] = (int64_t)H * W; dst[2] = W; dst[3] = 1;
  } else if (l == LAYOUT_NHWC) {
    dst[0] = (int64_t)C * H * W; dst[1] = 1; dst[2] = (int64_t)W * C; dst[3] = C;
  } else {
    for (int k = 0; k < 4; ++k) dst[k] = src[k];
  }
}

int fill_dims(Dims& d, int64_t N, int C, int H, int W, int64_t x_batch, int padding, int flags) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || N > 0x7fffffff) {
    set_error("invalid sizes N=%lld C=%d H=%d W=%d", (long long)N, C, H, W);
    return C2M_ERR_INVALID;
  }
  if (padding != C2M_PAD_BORDER && padding != C2M_PAD_ZEROS) {
    set_error("invalid padding %d", padding);
    return C2M_ERR_INVALID;
  }
  if (x_batch <= 0) x_batch = N;
  if (N > 0 && (x_batch > N || N % x_batch != 0)) {
    set_error("x_batch=%lld must divide N=%lld", (long long)x_batch, (long long)N);
    return C2M_ERR_INVALID;
  }
  d.N = (int)N; d.C = C; d.H = H; d.W = W;
  d.x_batch = (int)(x_batch > 0 ? x_batch : 1);
  d.padding = padding;
  // (align_corners=True lives in the stride-generic kernels only)
  d.flags = (flags & C2M_FLAG_ALIGN_CORNERS) ? (flags | C2M_FLAG_FORCE_GENERIC) : flags;
  // fp32 arithmetic exactly as torch.linspace / the reference's python scalars produce it
  d.stepx = 2.0f / (float)(W - 1);
  d.stepy = 2.0f / (float)(H - 1);
  d.bw = (float)((W - 1.0) / 2.0);
  d.bh = (float)((H - 1.0) / 2.0);
  d.inv_bw = 1.0f / d.bw;
  d.inv_bh = 1.0f / d.bh;
  return C2M_OK;
}

static int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

__global__ void base_grid_kernel(float* grid, int N, int H, int W, float stepx, float stepy) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = HW * N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / HW;
    const int r = (int)(idx - n * HW);
    const int i = r / W, j = r - i * W;
    grid[n * 2 * HW + r] = base_coord(j, W, stepx);
    grid[n * 2 * HW + HW + r] = base_coord(i, H, stepy);
  }
}

}  // namespace c2m

using namespace c2m;

extern "C" {

int c2m_warp_version(void) { return C2M_WARP_VERSION; }
const char* c2m_warp_last_error(void) { return g_err; }
uint64_t c2m_warp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int c2m_warp_profile(int enable) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (enable && !g_prof.a) {
    if (cudaEventCreate(&g_prof.a) != cudaSuccess || cudaEventCreate(&g_prof.b) != cudaSuccess) {
      set_error("c2m_warp_profile: cudaEventCreate failed");
      (void)cudaGetLastError();
      return C2M_ERR_CUDA;
    }
  }
  g_prof.on = enable != 0;
  g_prof.valid = false;
  if (!enable && g_prof.a) {
    // the library keeps no CUDA objects alive while the hook is off (nothing to outlive the caller's context)
    (void)cudaEventDestroy(g_prof.a);
    (void)cudaEventDestroy(g_prof.b);
    g_prof.a = g_prof.b = nullptr;
  }
  return C2M_OK;
}

float c2m_warp_profile_last_ms(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.a || !g_prof.valid) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_prof.b) != cudaSuccess || cudaEventElapsedTime(&ms, g_prof.a, g_prof.b) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1.f;
  }
  return ms;
}

static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

int c2m_warp_blend_fwd_rs(const float* x, const float* flow, const float* mask, const float* other, float* out,
                          int64_t N, int C, int H, int W, int64_t x_batch, const int64_t x_strides[4],
                          const int64_t out_strides[4], const c2m_resize* rs, int padding, int flags,
                          void* cuda_stream) {
  FwdParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_dims(p.d, N, C, H, W, x_batch, padding, flags);
  if (rc) return rc;
  if (N == 0 || C == 0 || H == 0 || W == 0) return C2M_OK;  // empty: nothing to write
  if ((rc = fill_resize(p.d, rs)) != C2M_OK) return rc;
  if (!x || !flow || !out || !x_strides || !out_strides) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if (other && !mask) {
    set_error("`other` requires a mask");
    return C2M_ERR_INVALID;
  }
  const Layout lx = classify(x_strides, p.d.x_batch, C, H, W), lo = classify(out_strides, N, C, H, W);
  canonical(p.xs, x_strides, lx, C, H, W);
  canonical(p.os, out_strides, lo, C, H, W);
  p.x = x; p.flow = flow; p.mask = mask; p.other = other; p.out = out;
  p.cchunk = C;
  rc = launch_fwd(p, lx, lo, reinterpret_cast<cudaStream_t>(cuda_stream));
  if (rc) return rc;
  return check_launch("c2m_warp_blend_fwd");
}

int c2m_warp_blend_fwd(const float* x, const float* flow, const float* mask, const float* other, float* out,
                       int64_t N, int C, int H, int W, int64_t x_batch, const int64_t x_strides[4],
                       const int64_t out_strides[4], int padding, int flags, void* cuda_stream) {
  return c2m_warp_blend_fwd_rs(x, flow, mask, other, out, N, C, H, W, x_batch, x_strides, out_strides, nullptr, padding,
                               flags, cuda_stream);
}

// fused-resize backward: [resized flow | resized mask | grad-flow at the feature size | grad-mask at the feature
// size] in front of the workspace of the plain call
static size_t resize_ws_bytes(int64_t N, int H, int W) {
  const size_t px = (size_t)N * H * W * sizeof(float);
  return 2 * up256(2 * px) + 2 * up256(px);
}

size_t c2m_warp_bwd_workspace_bytes_rs(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx,
                                       const c2m_resize* rs, int flags) {
  size_t b = bwd_workspace_bytes(N, C, H, W, x_batch, want_gx, flags);
  if (rs && N > 0 && H > 0 && W > 0) {
    const bool on = (rs->flow_h > 0 && rs->flow_h != H) || (rs->flow_w > 0 && rs->flow_w != W) ||
                    (rs->mask_h > 0 && rs->mask_h != H) || (rs->mask_w > 0 && rs->mask_w != W) || rs->fold_t > 0;
    if (on) b += resize_ws_bytes(N, H, W);
  }
  return b;
}

size_t c2m_warp_bwd_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx, int flags) {
  return bwd_workspace_bytes(N, C, H, W, x_batch, want_gx, flags);
}

int c2m_warp_blend_bwd_rs(const float* x, const float* flow, const float* mask, const float* other, const float* gout,
                          float* gx, float* gflow, float* gmask, float* gother, int64_t N, int C, int H, int W,
                          int64_t x_batch, const int64_t x_strides[4], const int64_t g_strides[4],
                          const c2m_resize* rs, int padding, int flags, void* workspace, size_t workspace_bytes,
                          void* cuda_stream) {
  BwdParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_dims(p.d, N, C, H, W, x_batch, padding, flags);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (N == 0 || H == 0 || W == 0) return C2M_OK;
  if ((rc = fill_resize(p.d, rs)) != C2M_OK) return rc;
  const Resize rsz = p.d.rs;
  const size_t gflow_elems = (size_t)N * 2 * ((rsz.on & 1) ? (size_t)rsz.fh * rsz.fw : (size_t)H * W);
  const size_t gmask_elems = (size_t)N * ((rsz.on & 2) ? (size_t)rsz.mh * rsz.mw : (size_t)H * W);
  if (C == 0) {  // no channels: flow / mask gradients are zero
    if ((gflow && cudaMemsetAsync(gflow, 0, gflow_elems * sizeof(float), st) != cudaSuccess) ||
        (gmask && cudaMemsetAsync(gmask, 0, gmask_elems * sizeof(float), st) != cudaSuccess)) {
      set_error("c2m_warp_blend_bwd: cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
      return C2M_ERR_CUDA;
    }
    return C2M_OK;
  }
  if (!x || !flow || !gout || !x_strides || !g_strides) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if ((other || gother) && !mask) {
    set_error("`other` requires a mask");
    return C2M_ERR_INVALID;
  }
  if (gmask && !mask) {
    set_error("gmask requested without a mask");
    return C2M_ERR_INVALID;
  }
  const Layout lx = classify(x_strides, p.d.x_batch, C, H, W), lg = classify(g_strides, N, C, H, W);
  if (gx && lx == LAYOUT_OTHER) {
    set_error("grad-input needs x to be NCHW- or NHWC-dense");
    return C2M_ERR_INVALID;
  }
  canonical(p.xs, x_strides, lx, C, H, W);
  canonical(p.gs, g_strides, lg, C, H, W);
  p.x = x; p.flow = flow; p.mask = mask; p.other = other; p.gout = gout;
  p.gx = gx; p.gflow = gflow; p.gmask = gmask; p.gother = gother;
  p.cchunk = C;
  if (rsz.on && (flags & C2M_FLAG_PLANNED)) {
    set_error("C2M_FLAG_PLANNED cannot be combined with a resized flow / mask");
    return C2M_ERR_INVALID;
  }
  if (rsz.on) {
    // the resized flow / mask are materialised once (the forward computes them on the fly), the kernels below run on
    // them unchanged, and the gradients go back through the resize in one gather pass
    const size_t front = resize_ws_bytes(N, H, W);
    if (!workspace || workspace_bytes < front + 256) {
      set_error("workspace too small: %zu < %zu", workspace_bytes, front + 256);
      return C2M_ERR_WORKSPACE;
    }
    const size_t px = (size_t)N * H * W * sizeof(float);
    char* b = reinterpret_cast<char*>(workspace);
    float* flow_s = reinterpret_cast<float*>(b);
    float* mask_s = reinterpret_cast<float*>(b + up256(2 * px));
    float* gflow_s = reinterpret_cast<float*>(b + up256(2 * px) + up256(px));
    float* gmask_s = reinterpret_cast<float*>(b + 2 * up256(2 * px) + up256(px));
    launch_resize_fwd(p.d, flow, mask, flow_s, mask_s, st);
    p.flow = flow_s;
    if (mask) p.mask = mask_s;
    if (gflow && (rsz.on & 1)) p.gflow = gflow_s;
    if (gmask && (rsz.on & 2)) p.gmask = gmask_s;
    p.d.rs.on = 0;
    rc = launch_bwd(p, lx, lg, b + front, workspace_bytes - front, st);
    if (rc) return rc;
    p.d.rs = rsz;
    launch_resize_bwd(p.d, gflow_s, gmask_s, gflow, gmask, st);
    return check_launch("c2m_warp_blend_bwd");
  }
  rc = launch_bwd(p, lx, lg, workspace, workspace_bytes, st);
  if (rc) return rc;
  return check_launch("c2m_warp_blend_bwd");
}

int c2m_warp_blend_bwd(const float* x, const float* flow, const float* mask, const float* other, const float* gout,
                       float* gx, float* gflow, float* gmask, float* gother, int64_t N, int C, int H, int W,
                       int64_t x_batch, const int64_t x_strides[4], const int64_t g_strides[4], int padding,
                       int flags, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  return c2m_warp_blend_bwd_rs(x, flow, mask, other, gout, gx, gflow, gmask, gother, N, C, H, W, x_batch, x_strides,
                               g_strides, nullptr, padding, flags, workspace, workspace_bytes, cuda_stream);
}

size_t c2m_warp_plan_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int flags) {
  Dims d;
  memset(&d, 0, sizeof(d));
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || N > 0x7fffffff) return 0;
  if (x_batch <= 0) x_batch = N;
  if (x_batch > N || N % x_batch != 0) return 0;
  d.N = (int)N; d.C = C; d.H = H; d.W = W; d.x_batch = (int)x_batch; d.flags = flags;
  return plan_bytes(d);
}

int c2m_warp_plan(const float* flow, const float* mask, int64_t N, int C, int H, int W, int64_t x_batch, int padding,
                  int flags, void* plan, size_t plan_bytes_, void* cuda_stream) {
  BwdParams p;
  memset(&p, 0, sizeof(p));
  const int rc = fill_dims(p.d, N, C, H, W, x_batch, padding, flags);
  if (rc) return rc;
  if (!flow) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  p.flow = flow;
  p.mask = mask;
  p.cchunk = C;
  const int64_t nhwc[4] = {(int64_t)C * H * W, 1, (int64_t)W * C, C};
  for (int k = 0; k < 4; ++k) p.xs[k] = p.gs[k] = nhwc[k];
  const int r2 = launch_plan(p, plan, plan_bytes_, reinterpret_cast<cudaStream_t>(cuda_stream));
  if (r2) return r2;
  return check_launch("c2m_warp_plan");
}

int c2m_relayout(const float* src, float* dst, int64_t N, int C, int H, int W, int to_channels_last,
                 void* cuda_stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || N > 65535 || (int64_t)H * W > 0x7fffffff) {
    set_error("c2m_relayout: invalid sizes");
    return C2M_ERR_INVALID;
  }
  if (N == 0 || C == 0 || H == 0 || W == 0) return C2M_OK;
  if (!src || !dst) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  launch_relayout(src, dst, N, C, H * W, to_channels_last != 0, reinterpret_cast<cudaStream_t>(cuda_stream));
  return check_launch("c2m_relayout");
}

int c2m_base_grid(float* grid, int64_t N, int H, int W, void* cuda_stream) {
  if (N < 0 || H < 0 || W < 0 || N > 0x7fffffff) {
    set_error("invalid sizes");
    return C2M_ERR_INVALID;
  }
  if (N == 0 || H == 0 || W == 0) return C2M_OK;
  if (!grid) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  const int64_t total = N * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  base_grid_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      grid, (int)N, H, W, 2.0f / (float)(W - 1), 2.0f / (float)(H - 1));
  count_launch();
  return check_launch("c2m_base_grid");
}

}  // extern "C"
