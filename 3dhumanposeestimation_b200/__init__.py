"""B200-native (sm_100a) implementation of the per-sample training / inference hot path of
AliEmreSenel/3DHumanPoseEstimation, behind the reference's own Python module API.

The package name starts with a digit, so import it with importlib (or through the root-level
``b200pose`` alias module)::

    import importlib
    pose = importlib.import_module("3dhumanposeestimation_b200")
    crit = pose.ComprehensivePoseLoss()

Host code is Python + PyTorch (device memory, streams, torch.distributed); every kernel on the hot
path lives in ``libpose_b200.so`` (hand-written CUDA, C ABI in ``include/pose_b200.h``).  There is no
CPU fallback and no multi-backend dispatch: CPU tensors raise.
"""
from .config import *  # noqa: F401,F403
from .model_config import ModelConfig  # noqa: F401
from .loss import ComprehensivePoseLoss  # noqa: F401
from .models.common import GaussianHeatmapGenerator, PoseRegressionHead  # noqa: F401
from .dataset.augmentation import PoseAugmentor  # noqa: F401
from .models.cnn import CNNPoseEstimation  # noqa: F401
from .models.transformers import TransformerPoseEstimation  # noqa: F401
from .optim import AdamW  # noqa: F401
from .dataset.transforms import Resize, resize_frames  # noqa: F401
from . import _lib, ops  # noqa: F401

__all__ = ["ModelConfig", "ComprehensivePoseLoss", "GaussianHeatmapGenerator", "PoseRegressionHead",
           "PoseAugmentor", "CNNPoseEstimation", "TransformerPoseEstimation", "AdamW", "Resize", "resize_frames"]
