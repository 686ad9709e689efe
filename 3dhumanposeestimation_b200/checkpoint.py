"""Checkpoint compatibility with the reference.

The reference saves ``{"step", "model_state_dict", "optimizer_state_dict", "model_args", "model_type"}`` with
``torch.save`` (src/train.py:300-309) and loads it in ``infer.load_pose_model`` (infer.py:73-131): a bare state dict
is accepted too, ``module.`` prefixes (DataParallel / DDP) are stripped, ``ModelConfig(model_type, **model_args)``
rebuilds the architecture, and a strict load falls back to ``strict=False``.  Both directions are reproduced here on
the drop-in modules, whose ``state_dict()`` keys and shapes equal the reference's, so checkpoints move freely between
the two code bases.
"""
from __future__ import annotations

import logging

import torch

from .model_config import ModelConfig

logger = logging.getLogger(__name__)


def build_model(model_type: str, config: ModelConfig):
    if model_type == "transformer":
        from .models.transformers import TransformerPoseEstimation
        return TransformerPoseEstimation(config)
    if model_type == "cnn":
        from .models.cnn import CNNPoseEstimation
        return CNNPoseEstimation(config)
    raise ValueError(f"Unknown model type: {model_type}")


def make_checkpoint(model, model_type: str, step: int = 0, optimizer=None) -> dict:
    """The dictionary the reference's training loop saves (src/train.py:300-306)."""
    return {
        "step": step,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
        "model_args": model.config.to_dict(),
        "model_type": model_type,
    }


def save_checkpoint(path: str, model, model_type: str, step: int = 0, optimizer=None) -> None:
    torch.save(make_checkpoint(model, model_type, step, optimizer), path)


def load_pose_model(checkpoint_path, model_type: str, device="cuda"):
    """infer.py:73-131: returns the model in eval mode on `device` (a reference checkpoint or one saved here)."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    state = None
    model_args = {}
    if isinstance(checkpoint, dict):
        model_type = checkpoint.get("model_type", model_type)
        model_args = dict(checkpoint.get("model_args", {}) or {})
        state = checkpoint.get("model_state_dict")
    if state is None:
        state = checkpoint          # the file IS the state dict
        if not isinstance(state, dict):
            raise ValueError("Checkpoint file does not appear to be a valid state_dict or contain a 'model_state_dict' key.")
    state = {k.replace("module.", ""): v for k, v in state.items()}
    if model_type == "transformer":
        model_args["vit_pretrained"] = False        # the weights come from the checkpoint, not from a timm download
    model = build_model(model_type, ModelConfig(model_type, **model_args))
    try:
        model.load_state_dict(state, strict=True)
    except RuntimeError as exc:
        logger.warning("Failed to load state_dict strictly (error: %s). Trying with strict=False.", exc)
        model.load_state_dict(state, strict=False)
    return model.to(device).eval()
