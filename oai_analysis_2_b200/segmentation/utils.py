"""Checkpoint loading for the prediction path (reference: oai_analysis/segmentation/utils.py:10-45)."""
import os

import torch


def load_checkpoint(path):
    """torch.load restricted to tensors, containers and numbers (weights_only=True): a checkpoint is data, not code.
    numpy scalars (e.g. a best_score saved as np.float64) are allow-listed."""
    import numpy as np
    safe = []
    core = getattr(np, "_core", None) or getattr(np, "core")
    for name in ("scalar", "_reconstruct"):
        if hasattr(core.multiarray, name):
            safe.append(getattr(core.multiarray, name))
    safe += [np.dtype, np.ndarray] + [type(np.dtype(t)) for t in ("float64", "float32", "int64", "int32", "bool")]
    with torch.serialization.safe_globals(safe):
        return torch.load(path, map_location="cpu", weights_only=True)


def initialize_model(model, optimizer=None, ckpoint_path=None):
    """Load checkpoint['model_state_dict'] (strict) or fall back to model.weights_init() when ckpoint_path is falsy.

    Returns (finished_epoch, best_score) like the reference.  A missing file raises ValueError (utils.py:41)."""
    finished_epoch, best_score = 0, 0
    if ckpoint_path:
        if not os.path.isfile(ckpoint_path):
            raise ValueError("=> no checkpoint found at '{}'".format(ckpoint_path))
        print("=> loading checkpoint '{}'".format(ckpoint_path))
        checkpoint = load_checkpoint(ckpoint_path)
        for key in ("best_score", "reg_best_score", "seg_best_score"):
            if key in checkpoint:
                best_score = checkpoint[key]
                break
        model.load_state_dict(checkpoint["model_state_dict"], strict=True)
        finished_epoch += checkpoint.get("epoch", 0)
        print("=> loaded checkpoint '{}' (epoch {})".format(ckpoint_path, checkpoint.get("epoch", 0)))
    else:
        model.weights_init()
    return finished_epoch, best_score
