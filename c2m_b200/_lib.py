"""ctypes binding of libc2m_warp.so (the C ABI declared in include/c2m_warp.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller
gets an exception -- never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes
import functools
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# C2M_WARP_LIB points at another build of the same library (kernel tuning: tools/build_variants.py)
LIB_PATH = os.environ.get("C2M_WARP_LIB") or os.path.join(_PKG, "libc2m_warp.so")

PAD_BORDER = 0
PAD_ZEROS = 1

FLAG_DETERMINISTIC = 0x1
FLAG_COORD_GRID = 0x2
FLAG_ALIGN_CORNERS = 0x4
FLAG_TRUE_DIV = 0x100
FLAG_NO_FMA = 0x200
FLAG_FORCE_GENERIC = 0x400
FLAG_NO_TMA = 0x800
FLAG_BWD_ATOMIC = 0x1000
FLAG_STAGE_NHWC = 0x4000
FLAG_NO_STAGE = 0x8000  # host-side only: keep NCHW tensors on the NCHW kernels (test / tuning hook)
FLAG_PLANNED = 0x20000  # bwd: the workspace is a plan made by c2m_warp_plan in the forward call
FLAG_STRICT_LAYOUT = 0x10000  # host-side only: results keep x's strides (NCHW x: staged backward, no promotion)

# every symbol include/c2m_warp.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = (
    "c2m_warp_version",
    "c2m_warp_last_error",
    "c2m_warp_blend_fwd",
    "c2m_warp_blend_bwd",
    "c2m_warp_bwd_workspace_bytes",
    "c2m_warp_blend_fwd_rs",
    "c2m_warp_blend_bwd_rs",
    "c2m_warp_bwd_workspace_bytes_rs",
    "c2m_base_grid",
    "c2m_relayout",
    "c2m_warp_plan_bytes",
    "c2m_warp_plan",
    "c2m_warp_launch_count",
    "c2m_occlusion_map",
    "c2m_occlusion_map_workspace_bytes",
    "c2m_warp_profile",
    "c2m_warp_profile_last_ms",
    "c2m_affine_warp",
    "c2m_sparse_motion",
    "c2m_warped_l1_workspace_bytes",
    "c2m_warped_l1_fwd",
    "c2m_warped_l1_bwd",
    "c2m_flow_consistency_workspace_bytes",
    "c2m_flow_consistency_fwd",
    "c2m_flow_consistency_bwd",
)

_lock = threading.Lock()
_lib = None

_i64 = ctypes.c_int64
_int = ctypes.c_int
_ptr = ctypes.c_void_p
_Strides = _i64 * 4

RESIZE_HALF_PIXEL = 0        # generator.py:84-85,91-92: align_corners=False, flow values not rescaled
RESIZE_CORNERS_RESCALE = 1   # utils.py:346-354: align_corners=True, flow values multiplied by new/old


class Resize(ctypes.Structure):
    """struct c2m_resize of include/c2m_warp.h: the sizes the flow / mask are passed at (0 = the feature size)."""
    _fields_ = [("flow_h", _int), ("flow_w", _int), ("mask_h", _int), ("mask_w", _int), ("flow_mode", _int),
                ("fold_t", _int)]


class C2MWarpError(RuntimeError):
    pass


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """dlopen the library (once). Raises C2MWarpError if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from . import _build
                _build.build()
            else:
                raise C2MWarpError(
                    f"{LIB_PATH} not found: build it with `python -m c2m_b200._build` "
                    "(there is no fallback path)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.c2m_warp_version.restype = _int
        lib.c2m_warp_version.argtypes = []
        lib.c2m_warp_last_error.restype = ctypes.c_char_p
        lib.c2m_warp_last_error.argtypes = []
        lib.c2m_warp_launch_count.restype = ctypes.c_uint64
        lib.c2m_warp_launch_count.argtypes = []
        lib.c2m_warp_blend_fwd.restype = _int
        lib.c2m_warp_blend_fwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _i64,
                                           ctypes.POINTER(_i64), ctypes.POINTER(_i64), _int, _int, _ptr]
        lib.c2m_warp_blend_bwd.restype = _int
        lib.c2m_warp_blend_bwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                           _i64, _int, _int, _int, _i64,
                                           ctypes.POINTER(_i64), ctypes.POINTER(_i64), _int, _int,
                                           _ptr, ctypes.c_size_t, _ptr]
        lib.c2m_warp_bwd_workspace_bytes.restype = ctypes.c_size_t
        lib.c2m_warp_bwd_workspace_bytes.argtypes = [_i64, _int, _int, _int, _i64, _int, _int]
        lib.c2m_warp_blend_fwd_rs.restype = _int
        lib.c2m_warp_blend_fwd_rs.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _i64,
                                              ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(Resize),
                                              _int, _int, _ptr]
        lib.c2m_warp_blend_bwd_rs.restype = _int
        lib.c2m_warp_blend_bwd_rs.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                              _i64, _int, _int, _int, _i64,
                                              ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(Resize),
                                              _int, _int, _ptr, ctypes.c_size_t, _ptr]
        lib.c2m_warp_bwd_workspace_bytes_rs.restype = ctypes.c_size_t
        lib.c2m_warp_bwd_workspace_bytes_rs.argtypes = [_i64, _int, _int, _int, _i64, _int, ctypes.POINTER(Resize), _int]
        lib.c2m_base_grid.restype = _int
        lib.c2m_base_grid.argtypes = [_ptr, _i64, _int, _int, _ptr]
        lib.c2m_warp_plan_bytes.restype = ctypes.c_size_t
        lib.c2m_warp_plan_bytes.argtypes = [_i64, _int, _int, _int, _i64, _int]
        lib.c2m_warp_plan.restype = _int
        lib.c2m_warp_plan.argtypes = [_ptr, _ptr, _i64, _int, _int, _int, _i64, _int, _int, _ptr, ctypes.c_size_t, _ptr]
        lib.c2m_relayout.restype = _int
        lib.c2m_relayout.argtypes = [_ptr, _ptr, _i64, _int, _int, _int, _int, _ptr]
        lib.c2m_warp_profile.restype = _int
        lib.c2m_warp_profile.argtypes = [_int]
        lib.c2m_warp_profile_last_ms.restype = ctypes.c_float
        lib.c2m_warp_profile_last_ms.argtypes = []
        lib.c2m_occlusion_map_workspace_bytes.restype = ctypes.c_size_t
        lib.c2m_occlusion_map_workspace_bytes.argtypes = [_i64, _int, _int]
        lib.c2m_occlusion_map.restype = _int
        lib.c2m_occlusion_map.argtypes = [_ptr, _ptr, _i64, _int, _int, _int, _ptr, ctypes.c_size_t, _ptr]
        lib.c2m_affine_warp.restype = _int
        lib.c2m_affine_warp.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _int, _int, _int, _ptr]
        lib.c2m_sparse_motion.restype = _int
        lib.c2m_sparse_motion.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _ptr]
        lib.c2m_warped_l1_workspace_bytes.restype = ctypes.c_size_t
        lib.c2m_warped_l1_workspace_bytes.argtypes = []
        lib.c2m_warped_l1_fwd.restype = _int
        lib.c2m_warped_l1_fwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _ptr, ctypes.c_size_t,
                                          _ptr]
        lib.c2m_warped_l1_bwd.restype = _int
        lib.c2m_warped_l1_bwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, _int, _ptr]
        lib.c2m_flow_consistency_workspace_bytes.restype = ctypes.c_size_t
        lib.c2m_flow_consistency_workspace_bytes.argtypes = [_i64, _int, _int, _int]
        lib.c2m_flow_consistency_fwd.restype = _int
        lib.c2m_flow_consistency_fwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _int, ctypes.c_float,
                                                 _ptr, ctypes.c_size_t, _ptr]
        lib.c2m_flow_consistency_bwd.restype = _int
        lib.c2m_flow_consistency_bwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int,
                                                 _int, ctypes.c_float, _ptr, ctypes.c_size_t, _ptr]
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().c2m_warp_last_error().decode("utf-8", "replace")
        raise C2MWarpError(f"{what} failed (code {rc}): {msg}")


@functools.lru_cache(maxsize=256)
def _strides_cached(s):
    return _Strides(*s)


def strides4(s) -> "_Strides":
    """ctypes int64[4] for a stride tuple (cached: the same few layouts come back on every call)."""
    return _strides_cached(tuple(int(v) for v in s))


@functools.lru_cache(maxsize=256)
def resize_spec(flow_h, flow_w, mask_h, mask_w, flow_mode, fold_t=0):
    """ctypes c2m_resize (cached)."""
    return Resize(int(flow_h), int(flow_w), int(mask_h), int(mask_w), int(flow_mode), int(fold_t))


def warp_blend_fwd(x_ptr, flow_ptr, mask_ptr, other_ptr, out_ptr, N, C, H, W, x_batch, x_strides, out_strides,
                   padding, flags, stream, resize=None) -> None:
    rc = load().c2m_warp_blend_fwd_rs(x_ptr, flow_ptr, mask_ptr, other_ptr, out_ptr, N, C, H, W, x_batch,
                                      strides4(x_strides), strides4(out_strides), resize, padding, flags, stream)
    _check(rc, "c2m_warp_blend_fwd")


def warp_blend_bwd(x_ptr, flow_ptr, mask_ptr, other_ptr, gout_ptr, gx_ptr, gflow_ptr, gmask_ptr, gother_ptr,
                   N, C, H, W, x_batch, x_strides, g_strides, padding, flags, ws_ptr, ws_bytes, stream,
                   resize=None) -> None:
    rc = load().c2m_warp_blend_bwd_rs(x_ptr, flow_ptr, mask_ptr, other_ptr, gout_ptr, gx_ptr, gflow_ptr, gmask_ptr,
                                      gother_ptr, N, C, H, W, x_batch, strides4(x_strides), strides4(g_strides),
                                      resize, padding, flags, ws_ptr, ws_bytes, stream)
    _check(rc, "c2m_warp_blend_bwd")


@functools.lru_cache(maxsize=1024)
def _ws_bytes_cached(N, C, H, W, x_batch, want_gx, flags, rs_key):
    rs = None if rs_key is None else resize_spec(*rs_key)
    return int(load().c2m_warp_bwd_workspace_bytes_rs(N, C, H, W, x_batch, want_gx, rs, flags))


def bwd_workspace_bytes(N, C, H, W, x_batch, want_gx, flags, resize=None) -> int:
    """Workspace size of the backward (cached per argument tuple: the query is a pure function of its arguments)."""
    key = None if resize is None else (resize.flow_h, resize.flow_w, resize.mask_h, resize.mask_w, resize.flow_mode,
                                       resize.fold_t)
    return _ws_bytes_cached(int(N), int(C), int(H), int(W), int(x_batch), int(bool(want_gx)), int(flags), key)


def base_grid(grid_ptr, N, H, W, stream) -> None:
    _check(load().c2m_base_grid(grid_ptr, N, H, W, stream), "c2m_base_grid")


@functools.lru_cache(maxsize=1024)
def plan_bytes(N, C, H, W, x_batch, flags) -> int:
    return int(load().c2m_warp_plan_bytes(N, C, H, W, x_batch, flags))


def warp_plan(flow_ptr, mask_ptr, N, C, H, W, x_batch, padding, flags, plan_ptr, nbytes, stream) -> None:
    _check(load().c2m_warp_plan(flow_ptr, mask_ptr, N, C, H, W, x_batch, padding, flags, plan_ptr, nbytes, stream),
           "c2m_warp_plan")


def relayout(src_ptr, dst_ptr, N, C, H, W, to_channels_last, stream) -> None:
    _check(load().c2m_relayout(src_ptr, dst_ptr, N, C, H, W, int(bool(to_channels_last)), stream), "c2m_relayout")


OCC_COORDS = 0x1
OCC_NO_CLAMP = 0x2


def occlusion_map(in_ptr, out_ptr, N, H, W, flags, ws_ptr, ws_bytes, stream) -> None:
    _check(load().c2m_occlusion_map(in_ptr, out_ptr, N, H, W, flags, ws_ptr, ws_bytes, stream), "c2m_occlusion_map")


def occlusion_map_workspace_bytes(N, H, W) -> int:
    return int(load().c2m_occlusion_map_workspace_bytes(N, H, W))


def affine_warp(theta_ptr, x_ptr, x_index_ptr, tx_ptr, flow_ptr, K, Kx, C, H, W, stream) -> None:
    _check(load().c2m_affine_warp(theta_ptr, x_ptr, x_index_ptr, tx_ptr, flow_ptr, K, Kx, C, H, W, stream),
           "c2m_affine_warp")


def sparse_motion(inst_ptr, ids_ptr, batch_ptr, thetas_ptr, bw_ptr, fw_ptr, bin_ptr, B, T, H, W, n_obj, stream) -> None:
    _check(load().c2m_sparse_motion(inst_ptr, ids_ptr, batch_ptr, thetas_ptr, bw_ptr, fw_ptr, bin_ptr, B, T, H, W,
                                    n_obj, stream), "c2m_sparse_motion")


def warped_l1_workspace_bytes() -> int:
    return int(load().c2m_warped_l1_workspace_bytes())


def warped_l1_fwd(src_ptr, flows_ptr, tgt_ptr, loss_ptr, B, C, T, H, W, ws_ptr, ws_bytes, stream) -> None:
    _check(load().c2m_warped_l1_fwd(src_ptr, flows_ptr, tgt_ptr, loss_ptr, B, C, T, H, W, ws_ptr, ws_bytes, stream),
           "c2m_warped_l1_fwd")


def warped_l1_bwd(src_ptr, flows_ptr, tgt_ptr, gloss_ptr, gflows_ptr, gtargets_ptr, B, C, T, H, W, stream) -> None:
    _check(load().c2m_warped_l1_bwd(src_ptr, flows_ptr, tgt_ptr, gloss_ptr, gflows_ptr, gtargets_ptr, B, C, T, H, W,
                                    stream), "c2m_warped_l1_bwd")


def flow_consistency_workspace_bytes(B, T, H, W) -> int:
    return int(load().c2m_flow_consistency_workspace_bytes(B, T, H, W))


def flow_consistency_fwd(flow_ptr, back_ptr, mfw_ptr, mbw_ptr, loss_ptr, B, T, H, W, scale, ws_ptr, ws_bytes,
                         stream) -> None:
    _check(load().c2m_flow_consistency_fwd(flow_ptr, back_ptr, mfw_ptr, mbw_ptr, loss_ptr, B, T, H, W, scale, ws_ptr,
                                           ws_bytes, stream), "c2m_flow_consistency_fwd")


def flow_consistency_bwd(flow_ptr, back_ptr, mfw_ptr, mbw_ptr, gloss_ptr, gflow_ptr, gback_ptr, gmfw_ptr, gmbw_ptr,
                         B, T, H, W, scale, ws_ptr, ws_bytes, stream) -> None:
    _check(load().c2m_flow_consistency_bwd(flow_ptr, back_ptr, mfw_ptr, mbw_ptr, gloss_ptr, gflow_ptr, gback_ptr,
                                           gmfw_ptr, gmbw_ptr, B, T, H, W, scale, ws_ptr, ws_bytes, stream),
           "c2m_flow_consistency_bwd")


def profile(enable: bool) -> None:
    """Measurement hook: bracket the dominant kernel of every call of this thread with CUDA events."""
    _check(load().c2m_warp_profile(int(bool(enable))), "c2m_warp_profile")


def profile_last_ms() -> float:
    return float(load().c2m_warp_profile_last_ms())


def launch_count() -> int:
    return int(load().c2m_warp_launch_count())
