// affine_warp.cu -- the second grid_sample flavour of the code base: the affine-grid object warp of
// src/modules/motion_estimator/dense_motion.py.
//
//   DenseMotionNetwork.warp (dense_motion.py:161-168)
//       grid = F.affine_grid(theta[1,2,3], x.size())            align_corners=False (default)
//       flow = (grid - base_grid) * ((w-1)/2, (h-1)/2)           base_grid = the reference's linspace(-1,1,.) grid
//       t_x  = F.grid_sample(x, grid)                            bilinear, ZEROS padding, align_corners=False
//   generate_sparse_motion (dense_motion.py:94-159) calls it objects x T times from a Python loop and merges the
//   results with torch.where(warped_obj == 1, ...) / torch.where(obj_mask == 1, ...) in object order.
//
// Kernels
//   affine_warp_kernel     K thetas in one launch (the batched `warp`): t_x [K,C,H,W] and flow [K,2,H,W].
//   sparse_motion_kernel   the whole loop of generate_sparse_motion in one launch: one thread per (b, t, i, j) walks
//                          the objects of image b in order, forms each object's mask on the fly from the instance map
//                          (never materialised), samples it at the affine grid location and applies the reference's
//                          `== 1` selections.  Writes sparse_motion_bw / _fw [B,2,T,H,W] and sparse_motion_bin [B,1,T,H,W].
//
// float32 arithmetic replicated bit for bit (settled on a B200 against torch 2.11, tools/probe_affine.py: 0 mismatches
// at six sizes): torch.linspace in its fma form (CPU == CUDA); base range linspace*(n-1)*(1/n) (ATen
// linspace_from_neg_one, align_corners=False); grid = fma(1, t2, fma(by, t1, bx*t0)) (the bmm's k-order); the sampler is
// make_geo's zeros-padding path, which reproduces ATen's (== 1) set exactly on binary masks.
#include <cstring>

#include "common.cuh"

namespace c2m {

struct AffineCoords {
  float lx, ly;  // linspace(-1,1,W)[j], linspace(-1,1,H)[i]  (the reference's base_grid)
  float bx, by;  // ATen's align_corners=False base grid
};

__device__ __forceinline__ AffineCoords affine_coords(const Dims& d, int i, int j, float rw, float rh) {
  AffineCoords c;
  c.lx = d.W > 1 ? base_coord(j, d.W, d.stepx) : -1.f;
  c.ly = d.H > 1 ? base_coord(i, d.H, d.stepy) : -1.f;
  // linspace_from_neg_one: num_steps <= 1 -> 0; else range * (n-1) / n (division by a scalar = reciprocal multiply)
  c.bx = d.W > 1 ? __fmul_rn(__fmul_rn(c.lx, (float)(d.W - 1)), rw) : 0.f;
  c.by = d.H > 1 ? __fmul_rn(__fmul_rn(c.ly, (float)(d.H - 1)), rh) : 0.f;
  return c;
}

// one row of base_grid[HW,3] @ theta^T: k = 0, 1, 2 accumulated in order with fused multiply-adds
__device__ __forceinline__ float affine_row(float bx, float by, float t0, float t1, float t2) {
  return fmaf(1.f, t2, fmaf(by, t1, __fmul_rn(bx, t0)));
}

template <bool WANT_FLOW>
__global__ void __launch_bounds__(256) affine_warp_kernel(const Dims d, const float* __restrict__ theta,
                                                          const float* __restrict__ x, const int* __restrict__ x_index,
                                                          int64_t Kx, float* __restrict__ t_x, float* __restrict__ flow,
                                                          float rw, float rh, float bwf, float bhf) {
  const int HW = d.H * d.W;
  const int64_t total = (int64_t)HW * d.N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / HW);
    const int r = (int)(idx - (int64_t)k * HW);
    const int i = r / d.W, j = r - i * d.W;
    const float* th = theta + (int64_t)k * 6;
    const AffineCoords c = affine_coords(d, i, j, rw, rh);
    const float gx = affine_row(c.bx, c.by, __ldg(th + 0), __ldg(th + 1), __ldg(th + 2));
    const float gy = affine_row(c.bx, c.by, __ldg(th + 3), __ldg(th + 4), __ldg(th + 5));
    if (WANT_FLOW) {
      flow[(int64_t)k * 2 * HW + r] = __fmul_rn(__fsub_rn(gx, c.lx), bwf);
      flow[(int64_t)k * 2 * HW + HW + r] = __fmul_rn(__fsub_rn(gy, c.ly), bhf);
    }
    if (t_x) {
      Geo g;
      make_geo<false, true>(d, gx, gy, i, j, g);  // COORD_GRID + zeros padding
      const int64_t img = x_index ? (int64_t)__ldg(x_index + k) : (int64_t)(k % Kx);
      const float* xb = x + img * d.C * HW;
      for (int ch = 0; ch < d.C; ++ch) {
        const float* xc = xb + (int64_t)ch * HW;
        const float vnw = g.oknw ? __ldg(xc + g.y0 * d.W + g.x0) : 0.f, vne = g.okne ? __ldg(xc + g.y0 * d.W + g.x1) : 0.f;
        const float vsw = g.oksw ? __ldg(xc + g.y1 * d.W + g.x0) : 0.f, vse = g.okse ? __ldg(xc + g.y1 * d.W + g.x1) : 0.f;
        float acc = vnw * g.wnw;
        acc = fmaf(vne, g.wne, acc);
        acc = fmaf(vsw, g.wsw, acc);
        acc = fmaf(vse, g.wse, acc);
        t_x[((int64_t)k * d.C + ch) * HW + r] = acc;
      }
    }
  }
}

struct SparseMotionParams {
  Dims d;  // N = B (images), C unused
  int T, n_obj;
  const float* instance;  // [B,1,H,W] instance ids as float
  const float* inst_ids;  // [n_obj] (0: skipped, dense_motion.py:126-127)
  const int* batch_ids;   // [n_obj]
  const float* thetas;    // [n_obj, T, 6]
  float* bw;              // [B,2,T,H,W] or NULL
  float* fw;              // [B,2,T,H,W] or NULL
  float* bin;             // [B,1,T,H,W] or NULL
  float rw, rh, bwf, bhf;
};

__global__ void __launch_bounds__(256) sparse_motion_kernel(const SparseMotionParams p) {
  extern __shared__ float s_obj[];  // [n_obj] ids, then [n_obj] batch ids (as int bits)
  const Dims& d = p.d;
  for (int k = threadIdx.x; k < p.n_obj; k += blockDim.x) {
    s_obj[k] = p.inst_ids[k];
    s_obj[p.n_obj + k] = __int_as_float(p.batch_ids[k]);
  }
  __syncthreads();
  const int HW = d.H * d.W;
  const int64_t total = (int64_t)d.N * p.T * HW;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % HW);
    const int bt = (int)(idx / HW);
    const int t = bt % p.T, b = bt / p.T;
    const int i = r / d.W, j = r - i * d.W;
    const AffineCoords c = affine_coords(d, i, j, p.rw, p.rh);
    const float* inst = p.instance + (int64_t)b * HW;
    const float own = __ldg(inst + r);
    float bwx = 0.f, bwy = 0.f, fwx = 0.f, fwy = 0.f, bin = 0.f;
    for (int k = 0; k < p.n_obj; ++k) {
      const float id = s_obj[k];
      if (__float_as_int(s_obj[p.n_obj + k]) != b || id == 0.f) continue;
      const float* th = p.thetas + ((int64_t)k * p.T + t) * 6;
      const float gx = affine_row(c.bx, c.by, __ldg(th + 0), __ldg(th + 1), __ldg(th + 2));
      const float gy = affine_row(c.bx, c.by, __ldg(th + 3), __ldg(th + 4), __ldg(th + 5));
      const float fx = __fmul_rn(__fsub_rn(gx, c.lx), p.bwf), fy = __fmul_rn(__fsub_rn(gy, c.ly), p.bhf);
      Geo g;
      make_geo<false, true>(d, gx, gy, i, j, g);
      // obj_mask = (instance == id).float(), sampled with zeros padding (ATen accumulation order nw, ne, sw, se)
      const float vnw = (g.oknw && __ldg(inst + g.y0 * d.W + g.x0) == id) ? 1.f : 0.f;
      const float vne = (g.okne && __ldg(inst + g.y0 * d.W + g.x1) == id) ? 1.f : 0.f;
      const float vsw = (g.oksw && __ldg(inst + g.y1 * d.W + g.x0) == id) ? 1.f : 0.f;
      const float vse = (g.okse && __ldg(inst + g.y1 * d.W + g.x1) == id) ? 1.f : 0.f;
      float acc = vnw * g.wnw;
      acc = fmaf(vne, g.wne, acc);
      acc = fmaf(vsw, g.wsw, acc);
      acc = fmaf(vse, g.wse, acc);
      if (acc == 1.f) {  // dense_motion.py:143-144,147-148: torch.where(warped_obj == 1, ...)
        bwx = fx;
        bwy = fy;
        bin = acc;
      }
      if (own == id) {  // dense_motion.py:145-146: torch.where(obj_mask == 1, obj_flow * -1, ...)
        fwx = -fx;
        fwy = -fy;
      }
    }
    const int64_t o2 = (((int64_t)b * 2) * p.T + t) * HW + r;  // [B,2,T,H,W]
    if (p.bw) {
      p.bw[o2] = bwx;
      p.bw[o2 + (int64_t)p.T * HW] = bwy;
    }
    if (p.fw) {
      p.fw[o2] = fwx;
      p.fw[o2 + (int64_t)p.T * HW] = fwy;
    }
    if (p.bin) p.bin[((int64_t)b * p.T + t) * HW + r] = bin;
  }
}

static int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

static int affine_dims(Dims& d, int64_t K, int C, int H, int W) {
  const int rc = fill_dims(d, K, C, H, W, K, C2M_PAD_ZEROS, C2M_FLAG_COORD_GRID);
  if (rc) return rc;
  if ((int64_t)H * W >= (1ll << 31)) {
    set_error("image too large");
    return C2M_ERR_INVALID;
  }
  return C2M_OK;
}

}  // namespace c2m

using namespace c2m;

extern "C" {

int c2m_affine_warp(const float* theta, const float* x, const int* x_index, float* t_x, float* flow, int64_t K,
                    int64_t Kx, int C, int H, int W, void* cuda_stream) {
  Dims d;
  memset(&d, 0, sizeof(d));
  int rc = affine_dims(d, K, C, H, W);
  if (rc) return rc;
  if (K == 0 || H == 0 || W == 0) return C2M_OK;
  if (!theta || (t_x && (!x || Kx <= 0 || C <= 0))) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  if (!t_x && !flow) return C2M_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const float rw = 1.0f / (float)W, rh = 1.0f / (float)H;
  const float bwf = (float)((W - 1.0) / 2.0), bhf = (float)((H - 1.0) / 2.0);  // tensor * python float (dense_motion.py:165)
  const int grid = grid_for(K * H * W);
  if (flow) affine_warp_kernel<true><<<grid, 256, 0, st>>>(d, theta, x, x_index, Kx > 0 ? Kx : 1, t_x, flow, rw, rh, bwf, bhf);
  else affine_warp_kernel<false><<<grid, 256, 0, st>>>(d, theta, x, x_index, Kx > 0 ? Kx : 1, t_x, flow, rw, rh, bwf, bhf);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_affine_warp: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

int c2m_sparse_motion(const float* instance, const float* inst_ids, const int* batch_ids, const float* thetas,
                      float* bw, float* fw, float* bin, int64_t B, int T, int H, int W, int n_obj, void* cuda_stream) {
  SparseMotionParams p;
  memset(&p, 0, sizeof(p));
  int rc = affine_dims(p.d, B, 1, H, W);
  if (rc) return rc;
  if (T < 0 || n_obj < 0 || n_obj > 4096) {
    set_error("invalid T=%d / n_obj=%d (at most 4096 objects)", T, n_obj);
    return C2M_ERR_INVALID;
  }
  if (B == 0 || T == 0 || H == 0 || W == 0) return C2M_OK;
  if (!instance || (n_obj > 0 && (!inst_ids || !batch_ids || !thetas))) {
    set_error("null pointer argument");
    return C2M_ERR_INVALID;
  }
  p.T = T; p.n_obj = n_obj;
  p.instance = instance; p.inst_ids = inst_ids; p.batch_ids = batch_ids; p.thetas = thetas;
  p.bw = bw; p.fw = fw; p.bin = bin;
  p.rw = 1.0f / (float)W; p.rh = 1.0f / (float)H;
  p.bwf = (float)((W - 1.0) / 2.0); p.bhf = (float)((H - 1.0) / 2.0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  sparse_motion_kernel<<<grid_for(B * T * H * W), 256, (size_t)n_obj * 2 * sizeof(float), st>>>(p);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("c2m_sparse_motion: %s", cudaGetErrorString(e));
    return C2M_ERR_CUDA;
  }
  return C2M_OK;
}

}  // extern "C"
