"""Checkpoint loading for the prediction path (reference: oai_analysis/segmentation/utils.py:10-45)."""
import os

import torch


def initialize_model(model, optimizer=None, ckpoint_path=None):
    """Load checkpoint['model_state_dict'] (strict) or fall back to model.weights_init() when ckpoint_path is falsy.

    Returns (finished_epoch, best_score) like the reference.  A missing file raises ValueError (utils.py:41)."""
    finished_epoch, best_score = 0, 0
    if ckpoint_path:
        if not os.path.isfile(ckpoint_path):
            raise ValueError("=> no checkpoint found at '{}'".format(ckpoint_path))
        print("=> loading checkpoint '{}'".format(ckpoint_path))
        checkpoint = torch.load(ckpoint_path, map_location="cpu", weights_only=False)
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
