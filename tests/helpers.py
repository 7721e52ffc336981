"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def dice(a, b):
    a, b = np.asarray(a) > 0.5, np.asarray(b) > 0.5
    s = a.sum() + b.sum()
    return 1.0 if s == 0 else 2.0 * (a & b).sum() / s


def write_seg_config(tmpdir, sd, patch, bias, BN, overlap, device="cuda", n_classes=2):
    """Checkpoint + training-config JSON + segmenter_config dict in the reference's formats
    (analysis_object.py:18-26, segmenter.py:52-56, utils.py:20-41)."""
    ck = os.path.join(str(tmpdir), "segmentation_model.pth.tar")
    torch.save({"model_state_dict": sd, "epoch": 600, "best_score": 0.0}, ck)
    cfg = os.path.join(str(tmpdir), "segmentation_train_config.pth.tar")
    with open(cfg, "w") as f:
        json.dump({"patch_size": list(patch), "model": "UNet",
                   "model_setting": {"in_channels": 1, "n_classes": n_classes, "bias": bias, "BN": BN}}, f)
    return dict(ckpoint_path=ck, training_config_file=cfg, device=device, batch_size=4, overlap_size=tuple(overlap),
                output_prob=True, output_itk=True)
