/*
 * c2m_warp.h -- C ABI of libc2m_warp.so: the fused flow-warp + occlusion-blend op of C2M for
 * NVIDIA B200 (sm_100a).  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * What each entry point replaces in the reference (PierfrancescoArdino/C2M):
 *
 *   c2m_warp_blend_fwd   src/utils/ops.py:187-193 `resample` (base grid :196-202 + flow
 *                        normalisation :190 + F.grid_sample bilinear/border :183-184) followed by
 *                        `* occlusion_map` at src/modules/generator/generator.py:93 and
 *                        src/modules/motion_estimator/motion_autoencoder.py:125.
 *   c2m_warp_blend_bwd   the autograd graph of the above (ATen grid_sampler_2d_backward, mul
 *                        backward, div/cat/add backward) -- the reference has no source for it.
 *   c2m_base_grid        src/utils/ops.py:196-202 `get_grid` (device-side, bit-identical to the
 *                        CPU-built float32 linspace grid).
 *   c2m_occlusion_map    src/utils/ops.py:205-275 `get_corresponding_map` / `get_occlusion_map` (the forward
 *                        splat that produces the warp's mask, dense_motion.py:148,151).
 *   c2m_affine_warp / c2m_sparse_motion  src/modules/motion_estimator/dense_motion.py:161-168 (`warp`: affine_grid +
 *                        zeros-padding grid_sample) and :94-152 (the objects x T Python loop around it).
 *   c2m_warped_l1_fwd/bwd  src/losses/losses.py:219-222: the T `resample` calls on the source frame, their
 *                        torch.cat and L1MaskedLoss (losses.py:184-189, no mask), and its autograd.
 *   c2m_flow_consistency_fwd/bwd  src/losses/losses.py:115-141 `FlowConsistLoss` (two C = 2 `resample` calls, abs,
 *                        optional masks, means) and its autograd.
 *
 * The reference's own native-operator convention (its only hand-written warp,
 * src/modules/third_party/resample2d/src/resample2d_cuda.cc:6-33) is followed where it makes
 * sense: the caller allocates every output, tensors are borrowed for the duration of the call,
 * work is enqueued on the stream that is passed in and the call returns without synchronising.
 * Unlike that operator, errors are reported (return code + c2m_warp_last_error()).
 *
 * Conventions
 *   - all tensors are float32 device memory on the current CUDA device;
 *   - x / out / gout / gx / other are logical [N, C, H, W] with element strides passed explicitly
 *     (NCHW-contiguous and channels-last have dedicated kernels, any other stride pattern runs a
 *     generic kernel); flow is [N, 2, H, W] contiguous, channel 0 = x displacement in pixels;
 *     mask is [N, 1, H, W] contiguous or NULL;
 *   - `other` NULL  => out = mask * warp(x)                       (reference semantics)
 *     `other` given => out = mask * warp(x) + (1 - mask) * other  (north-star blend, needs mask); `other` has the
 *     output's strides (out_strides in the forward, g_strides in the backward, like gother);
 *   - x_batch: number of distinct images in x. x_batch == N (or 0) is the plain case. x_batch < N
 *     (N % x_batch == 0) means frame n reads image n % x_batch, i.e. the T-fold repeat of
 *     motion_autoencoder.py:117-119 without materialising the copies; gx is then [x_batch,C,H,W]
 *     and receives the sum over the repeats;
 *   - return 0 on success, a C2M_ERR_* code otherwise (message: c2m_warp_last_error(), thread
 *     local).  No hidden synchronisation, no allocation, no global mutable state besides a
 *     once-initialised driver entry point.
 */
#ifndef C2M_WARP_H_
#define C2M_WARP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define C2M_API __attribute__((visibility("default")))
#else
#define C2M_API
#endif

#define C2M_WARP_VERSION 210

/* padding (ATen GridSamplerPadding): the reference path uses border (ops.py:184); zeros is the
 * flavour of src/modules/motion_estimator/dense_motion.py:167 */
#define C2M_PAD_BORDER 0
#define C2M_PAD_ZEROS 1

/* flags */
#define C2M_FLAG_DETERMINISTIC 0x1  /* bwd: bitwise run-to-run reproducible grad-input */
#define C2M_FLAG_COORD_GRID 0x2     /* `flow` is a normalised sampling grid [N,H,W,2] (utils.grid_sample, ops.py:183); gflow has that shape too */
#define C2M_FLAG_ALIGN_CORNERS 0x4  /* sample with align_corners=True: ix = (c + 1) / 2 * (W - 1).  No call site of the
                                       reference asks for it (its base grid is built for this convention but sampled
                                       with the other); served by the stride-generic kernels */
#define C2M_FLAG_TRUE_DIV 0x100     /* probe only: divide by (size-1)/2 (ATen CPU) instead of * reciprocal (ATen CUDA) */
#define C2M_FLAG_NO_FMA 0x200       /* probe only: unfused (c+1)*size-1 */
#define C2M_FLAG_FORCE_GENERIC 0x400 /* run the stride-generic kernels (test hook) */
#define C2M_FLAG_NO_TMA 0x800       /* stage flow/mask with plain loads instead of TMA (test hook) */
#define C2M_FLAG_BWD_ATOMIC 0x1000  /* bwd: force the direct global-atomics scatter (test hook) */
#define C2M_FLAG_STAGE_NHWC 0x4000  /* bwd, NCHW-dense tensors: stage gout / x / gx through channels-last copies in the
                                       workspace and run the channels-last kernels (pass the flag to the workspace
                                       query too: three tensor-sized staging buffers are added) */

#define C2M_OK 0
#define C2M_ERR_INVALID 1   /* bad argument (null pointer, size, stride, alignment) */
#define C2M_ERR_CUDA 2      /* a CUDA runtime / driver call failed */
#define C2M_ERR_WORKSPACE 3 /* workspace too small */

C2M_API int c2m_warp_version(void);
C2M_API const char* c2m_warp_last_error(void);

C2M_API int c2m_warp_blend_fwd(const float* x, const float* flow, const float* mask, const float* other,
                       float* out, int64_t N, int C, int H, int W, int64_t x_batch,
                       const int64_t x_strides[4], const int64_t out_strides[4],
                       int padding, int flags, void* cuda_stream);

/* Any of gx / gflow / gmask / gother may be NULL (<=> ctx.needs_input_grad false).  gx is fully
 * written by the call (zero-filled first when the scatter uses atomics); it has x's strides.
 * gout and gother have g_strides.  workspace: c2m_warp_bwd_workspace_bytes() bytes for the SAME
 * flags AND the same N, C, H, W, x_batch (the size is not monotonic in N: small many-channel levels get extra
 * room for the channel-sliced kernels), 256-byte aligned, contents undefined on entry and exit (the query is not
 * told the layout and sizes for the larger of the channels-last and NCHW schemes). */
C2M_API int c2m_warp_blend_bwd(const float* x, const float* flow, const float* mask, const float* other,
                       const float* gout, float* gx, float* gflow, float* gmask, float* gother,
                       int64_t N, int C, int H, int W, int64_t x_batch,
                       const int64_t x_strides[4], const int64_t g_strides[4],
                       int padding, int flags, void* workspace, size_t workspace_bytes,
                       void* cuda_stream);

C2M_API size_t c2m_warp_bwd_workspace_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx,
                                    int flags);

/* The same op with the bilinear resize of the flow and the mask fused in.  The reference resizes
 * both to the feature size right before every warp:
 *   C2M_RESIZE_HALF_PIXEL       src/modules/generator/generator.py:84-85,91-92 -- F.interpolate(bilinear), i.e.
 *                               align_corners=False, flow VALUES NOT rescaled;
 *   C2M_RESIZE_CORNERS_RESCALE  src/utils/utils.py:346-354 `resize_flow` (+ motion_autoencoder.py:120-124) --
 *                               align_corners=True, flow x / y values divided by old_w/new_w and old_h/new_h.
 * The mask is always resized with align_corners=False (generator.py:92, motion_autoencoder.py:122-124).
 * `flow` is [N,2,flow_h,flow_w], `mask` [N,1,mask_h,mask_w] (0 = the feature size H, W: no resize of that tensor).
 * Forward: one launch, the taps are gathered inside the warp kernel.  Backward: gflow / gmask have the shapes of
 * flow / mask as passed (the gradient is taken through the resize: ATen upsample_bilinear2d_backward + the divides),
 * written in full, bitwise reproducible (a gather over source pixels, no atomics).  rs == NULL: same as the calls above. */
#define C2M_RESIZE_HALF_PIXEL 0
#define C2M_RESIZE_CORNERS_RESCALE 1
typedef struct c2m_resize {
  int flow_h, flow_w; /* size of `flow` as passed; 0 = H, W */
  int mask_h, mask_w; /* size of `mask` as passed; 0 = H, W */
  int flow_mode;      /* C2M_RESIZE_* (ignored when the flow is not resized) */
  int fold_t;         /* 0: flow [N,2,h,w], mask [N,1,h,w].  T > 0: the tensors are the reference's 5-D clips, flow
                         [B,2,T,h,w] and mask [B,1,T,h,w] with B = N / T, and frame n = t * B + b reads plane (b, :, t)
                         -- the fold `torch.cat(torch.unbind(., 2), 0)` of motion_autoencoder.py:120-123 without the
                         copies; gflow / gmask are written in the same 5-D layout */
} c2m_resize;

C2M_API int c2m_warp_blend_fwd_rs(const float* x, const float* flow, const float* mask, const float* other,
                                  float* out, int64_t N, int C, int H, int W, int64_t x_batch,
                                  const int64_t x_strides[4], const int64_t out_strides[4], const c2m_resize* rs,
                                  int padding, int flags, void* cuda_stream);
C2M_API int c2m_warp_blend_bwd_rs(const float* x, const float* flow, const float* mask, const float* other,
                                  const float* gout, float* gx, float* gflow, float* gmask, float* gother,
                                  int64_t N, int C, int H, int W, int64_t x_batch,
                                  const int64_t x_strides[4], const int64_t g_strides[4], const c2m_resize* rs,
                                  int padding, int flags, void* workspace, size_t workspace_bytes,
                                  void* cuda_stream);
C2M_API size_t c2m_warp_bwd_workspace_bytes_rs(int64_t N, int C, int H, int W, int64_t x_batch, int want_gx,
                                               const c2m_resize* rs, int flags);

/* [N,2,H,W] base grid, bit-identical to the reference's CPU float32 construction. */
C2M_API int c2m_base_grid(float* grid, int64_t N, int H, int W, void* cuda_stream);

/* Plan of the channels-last float backward.  The first stage of c2m_warp_blend_bwd -- registering every row segment
 * of 32 output pixels with the destination tiles its samples touch -- depends on the flow and the mask only.  A
 * caller that knows at forward time that grad-input will be asked for can run it then, on a second stream next to
 * the forward kernel (that stage is bound by integer instructions, the forward by HBM), into a buffer of
 * c2m_warp_plan_bytes() bytes, and later pass that buffer as the `workspace` of c2m_warp_blend_bwd together with
 * C2M_FLAG_PLANNED: the backward then starts with its gather kernel.  c2m_warp_plan_bytes() == 0: this
 * configuration has no plan (deterministic mode, test-hook flags, sizes beyond the gather's 32-bit limits).  A plan is
 * consumed by the backward that uses it; x, gout and gx of that backward must be channels-last dense and 16-byte
 * aligned, the flow / mask not resized (C2M_ERR_INVALID otherwise). */
#define C2M_FLAG_PLANNED 0x20000
C2M_API size_t c2m_warp_plan_bytes(int64_t N, int C, int H, int W, int64_t x_batch, int flags);
C2M_API int c2m_warp_plan(const float* flow, const float* mask, int64_t N, int C, int H, int W, int64_t x_batch,
                          int padding, int flags, void* plan, size_t plan_bytes, void* cuda_stream);
/* Layout change of a dense [N,C,H,W] float32 tensor: NCHW-contiguous -> channels-last (to_channels_last != 0) or
 * back.  The reference's tensors are NCHW (its convolutions produce them so); the channels-last kernels are the
 * fast ones on B200, so the Python host converts an NCHW `x` ONCE in the forward, keeps that copy for the backward
 * and returns channels-last strided results (the memory format then propagates through the caller's
 * convolutions like any channels_last tensor in PyTorch) instead of staging three tensors in every backward. */
C2M_API int c2m_relayout(const float* src, float* dst, int64_t N, int C, int H, int W, int to_channels_last,
                         void* cuda_stream);

/* Forward-splat occlusion map, the producer of the warp's mask (reference src/utils/ops.py:263-275
 * `get_occlusion_map`; with C2M_OCC_COORDS | C2M_OCC_NO_CLAMP: ops.py:205-251 `get_corresponding_map`).
 * in [N,2,H,W] contiguous (flow in pixels, or absolute coordinates), out [N,1,H,W].  The sum is accumulated in
 * 64-bit fixed point: bitwise reproducible, unlike the reference's scatter_add_.  workspace:
 * c2m_occlusion_map_workspace_bytes() bytes. */
#define C2M_OCC_COORDS 0x1   /* `in` holds absolute pixel coordinates instead of a flow */
#define C2M_OCC_NO_CLAMP 0x2 /* do not clamp the sum to [0, 1] */
C2M_API size_t c2m_occlusion_map_workspace_bytes(int64_t N, int H, int W);
C2M_API int c2m_occlusion_map(const float* in, float* out, int64_t N, int H, int W, int flags, void* workspace,
                              size_t workspace_bytes, void* cuda_stream);

/* Affine-grid object warp, batched (reference src/modules/motion_estimator/dense_motion.py:161-168
 * `DenseMotionNetwork.warp`: F.affine_grid(theta) with align_corners=False, flow = (grid - base_grid) * ((w-1)/2, (h-1)/2),
 * t_x = F.grid_sample(x, grid) bilinear / ZEROS padding -- called objects x T times from the Python loop at :123-142).
 * theta [K,2,3]; x [Kx,C,H,W] contiguous; x_index [K] int32 (theta k samples image x_index[k]) or NULL (image k % Kx);
 * t_x [K,C,H,W] and flow [K,2,H,W], either may be NULL.  One launch for all K; forward only (the reference detaches the
 * flows, dense_motion.py:149-151). */
C2M_API int c2m_affine_warp(const float* theta, const float* x, const int* x_index, float* t_x, float* flow,
                            int64_t K, int64_t Kx, int C, int H, int W, void* cuda_stream);

/* The whole object loop of `generate_sparse_motion` (dense_motion.py:94-152) in one launch.  instance [B,1,H,W]: the
 * instance map as float; inst_ids [n_obj] float (0 = skipped, :126-127), batch_ids [n_obj] int32, thetas [n_obj,T,2,3].
 * Per pixel the objects of its image are visited in index order (later objects overwrite earlier ones, as the loop does):
 *   warped = grid_sample((instance == id).float(), affine_grid(theta))      the object mask is never materialised
 *   bw  [B,2,T,H,W] = obj_flow where warped == 1        (:143-144)
 *   fw  [B,2,T,H,W] = -obj_flow where instance == id    (:145-146)
 *   bin [B,1,T,H,W] = warped where warped == 1          (:147-148)
 * zero elsewhere; any output may be NULL.  At most 4096 objects. */
C2M_API int c2m_sparse_motion(const float* instance, const float* inst_ids, const int* batch_ids, const float* thetas,
                              float* bw, float* fw, float* bin, int64_t B, int T, int H, int W, int n_obj,
                              void* cuda_stream);

/* Fused warped-frame L1 loss (reference src/losses/losses.py:219-222: T calls of utils.resample on the source frame,
 * torch.cat, then L1MaskedLoss without a mask, losses.py:184-189):
 *   loss = mean over (b,c,t,i,j) of | resample(source, flows[:,:,t])[b,c,i,j] - targets[b,c,t,i,j] |
 * source [B,C,H,W], flows [B,2,T,H,W] (pixels, channel 0 = x), targets [B,C,T,H,W], all contiguous float32 device
 * memory; loss / gloss: one float in device memory.  The sum is formed from per-block partial sums in double, in a
 * fixed order (bitwise reproducible).  Backward writes d loss / d flows (all of it) and, when asked, d loss / d targets;
 * the source frame is data in the reference and gets no gradient here.  workspace: c2m_warped_l1_workspace_bytes(). */
C2M_API size_t c2m_warped_l1_workspace_bytes(void);
C2M_API int c2m_warped_l1_fwd(const float* source, const float* flows, const float* targets, float* loss, int64_t B,
                              int C, int T, int H, int W, void* workspace, size_t workspace_bytes, void* cuda_stream);
C2M_API int c2m_warped_l1_bwd(const float* source, const float* flows, const float* targets, const float* gloss,
                              float* gflows /* nullable */, float* gtargets /* nullable */, int64_t B, int C, int T,
                              int H, int W, void* cuda_stream);

/* Fused forward-backward flow-consistency loss (reference src/losses/losses.py:115-141 `FlowConsistLoss`):
 *   loss = scale * ( mean(mask_bw * |resample(flow, flowback) + flowback|) + mean(mask_fw * |resample(flowback, flow) + flow|) )
 * flow / flowback [B,2,T,H,W] (pixels, channel 0 = x), mask_fw / mask_bw [B,1,T,H,W] or both NULL, all contiguous
 * float32; scale = num_predicted_frames (losses.py:141); loss / gloss: one float in device memory.  The 5-D tensors
 * are read in place (the reference folds the frame axis into the batch with four torch.cat copies first).  Sums in
 * double in a fixed order; the scatter part of the backward in 64-bit fixed point: bitwise reproducible.  Any gradient
 * pointer may be NULL.  workspace: c2m_flow_consistency_workspace_bytes() for the same sizes. */
C2M_API size_t c2m_flow_consistency_workspace_bytes(int64_t B, int T, int H, int W);
C2M_API int c2m_flow_consistency_fwd(const float* flow, const float* flowback, const float* mask_fw,
                                     const float* mask_bw, float* loss, int64_t B, int T, int H, int W, float scale,
                                     void* workspace, size_t workspace_bytes, void* cuda_stream);
C2M_API int c2m_flow_consistency_bwd(const float* flow, const float* flowback, const float* mask_fw,
                                     const float* mask_bw, const float* gloss, float* gflow, float* gflowback,
                                     float* gmask_fw, float* gmask_bw, int64_t B, int T, int H, int W, float scale,
                                     void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Measurement hook (process wide, off by default, not meant for concurrent callers).  While enabled, every
 * forward / backward call brackets its
 * dominant kernel -- the fused forward kernel; the gather kernel of the backward -- with a pair of CUDA events on the
 * call's stream.  c2m_warp_profile_last_ms() waits for the stop event of the most recent call and
 * returns that kernel's duration in milliseconds (< 0: nothing recorded).  Used by bench.py for `roofline`. */
C2M_API int c2m_warp_profile(int enable);
C2M_API float c2m_warp_profile_last_ms(void);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
C2M_API uint64_t c2m_warp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* C2M_WARP_H_ */
