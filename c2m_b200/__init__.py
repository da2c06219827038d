"""c2m_b200 -- B200-native (sm_100a) implementation of C2M's flow-guided feature warping hot path.

Only what that path needs lives here:
  csrc/          hand-written CUDA kernels + the C ABI (libc2m_warp.so, include/c2m_warp.h)
  _lib           ctypes binding (no fallback: a missing library raises)
  functional     torch.autograd.Function over the C ABI
  ops            resample / grid_sample / get_grid   (reference src/utils/ops.py:183-202),
                 get_occlusion_map / get_corresponding_map   (ops.py:205-275)
  loss           warped_l1_loss   (the `warped` term of the training losses, src/losses/losses.py:219-222)
  loss           flow_consistency_loss / FlowConsistLoss   (losses.py:115-141)
  motion         affine_warp / sparse_motion / generate_sparse_motion   (the affine-grid object warp and its
                 objects x T loop, src/modules/motion_estimator/dense_motion.py:94-168)
  generator      deform_input / apply_optical / decoder_warp / patch_reference
                 (reference src/modules/generator/generator.py:80-96,
                  src/modules/motion_estimator/motion_autoencoder.py:117-125)
"""
from .functional import WarpBlendFunction, warp_blend  # noqa: F401
from .generator import apply_optical, decoder_warp, deform_input, patch_reference, resize_flow  # noqa: F401
from .loss import FlowConsistLoss, WarpedL1Function, flow_consistency_loss, warped_l1_loss  # noqa: F401
from .motion import affine_warp, generate_sparse_motion, sparse_motion  # noqa: F401
from .ops import get_corresponding_map, get_grid, get_occlusion_map, grid_sample, resample  # noqa: F401

__version__ = "0.1.0"
