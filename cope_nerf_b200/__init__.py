"""cope_nerf_b200 — B200-native (sm_100a) implementation of cope-nerf's NeuS render + train hot path behind the
reference's own module API (model/neus_fields.py, model/neus_renderer.py, model/neus_embedder.py,
model/poses_retriever.py, model/common.py, model/training.py).

    from cope_nerf_b200 import SDFNetwork, RenderingNetwork, SingleVarianceNetwork, NeuSRenderer, PoseRetriever, MotionNetwork

The compute path is hand-written CUDA in libcope_b200.so (include/cope_b200.h).  There is no CPU or PyTorch
fallback: kernels raise `CopeError` when the library or a CUDA device is missing.
"""
from ._lib import CopeError, PREC_BF16, PREC_FP32, load as load_library
from .embedder import get_embedder
from .fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
from .renderer import NeuSRenderer, sample_pdf
from .motion import MotionNetwork
from .common import (Exp, PoseRetriever, arange_pixels, convert3x4_4x4, get_world_cameraOrigin_cameraRay, make_c2w,
                     pixels_from_indices, vec2skew)
from . import losses, optim, pose_refinement, training

# precision bench.py / smoke use when none is requested: the tensor-core path (strict fp32 parity mode: PREC_FP32)
DEFAULT_PRECISION = PREC_BF16

__all__ = ["CopeError", "PREC_BF16", "PREC_FP32", "load_library", "get_embedder", "RenderingNetwork", "SDFNetwork", "MotionNetwork",
           "SingleVarianceNetwork", "NeuSRenderer", "sample_pdf", "Exp", "PoseRetriever", "arange_pixels",
           "convert3x4_4x4", "get_world_cameraOrigin_cameraRay", "make_c2w", "pixels_from_indices", "vec2skew",
           "training", "losses", "pose_refinement", "optim"]
