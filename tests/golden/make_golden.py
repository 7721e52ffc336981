"""Generate the segmentation golden fixtures by running the REFERENCE'S OWN code (read-only /root/reference).

Run once in the build container:   python tests/golden/make_golden.py
It cannot run on the GPU box (no /root/reference there); the .npz files it writes are committed.

The reference's prediction path touches `itk` only through GetArrayFromImage / GetImageFromArray / CopyInformation
(oai_analysis/segmentation/image_transforms.py:403,516-517), so a 10-line stub module stands in for the missing
ITK wheel; everything else -- Partition, UNet, initialize_model, Segmenter3DInPatchClassWise.segment -- is the
reference, unmodified.  Weights come from oracle.seg_oracle.make_unet_state_dict (numpy PCG64, version-stable) and
are loaded through the reference's own checkpoint loader (segmentation/utils.py:20-41).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.seg_oracle import make_unet_state_dict, synthetic_knee  # noqa: E402

CASES = {
    # name: volume zyx, patch xyz, overlap xyz, BN, bias, seed, head calibration (oracle.seg_oracle.calibrate_head)
    "seg_small_pertap": dict(shape=(12, 100, 50), patch=[64, 128, 16], overlap=(8, 16, 4), BN=True, bias=True, seed=11,
                             head_gain=17.080657075570716, head_bias=[1.7763312324057952, -2.209280787428821]),
    "seg_small_nobn": dict(shape=(12, 100, 50), patch=[64, 128, 16], overlap=(8, 16, 4), BN=False, bias=True, seed=12,
                           head_gain=112.94203260602143, head_bias=[2.5398848029828915, -1.158113843700921]),
    "seg_prod_tile": dict(shape=(20, 100, 100), patch=[128, 128, 32], overlap=(16, 16, 8), BN=True, bias=True,
                          seed=13, head_gain=88.16291826695385, head_bias=[-4.068152692193752, -7.490846258792958]),
}


def install_itk_stub():
    class _Img:
        def __init__(self, arr):
            self.arr = arr

        def CopyInformation(self, other):
            self.info_from = other

    stub = types.ModuleType("itk")
    stub.GetArrayFromImage = lambda im: np.asarray(im.arr if isinstance(im, _Img) else im)
    stub.GetImageFromArray = lambda a: _Img(a)
    sys.modules["itk"] = stub
    return _Img


def run_reference(case, tmp):
    sys.path.insert(0, "/root/reference")
    from oai_analysis.segmentation.segmenter import Segmenter3DInPatchClassWise

    sd = make_unet_state_dict(case["seed"], 1, 2, case["bias"], case["BN"], True, case["head_gain"], case["head_bias"])
    ck = os.path.join(tmp, "ck.pth.tar")
    torch.save({"model_state_dict": sd, "epoch": 0, "best_score": 0.0}, ck)
    cfg_json = os.path.join(tmp, "train_config.json")
    with open(cfg_json, "w") as f:
        json.dump({"patch_size": case["patch"], "model": "UNet",
                   "model_setting": {"in_channels": 1, "n_classes": 2, "bias": case["bias"], "BN": case["BN"]}}, f)
    config = dict(ckpoint_path=ck, training_config_file=cfg_json, device="cpu", batch_size=4,
                  overlap_size=case["overlap"], output_prob=True, output_itk=True)
    seg = Segmenter3DInPatchClassWise(mode="pred", config=config)
    vol = synthetic_knee(case["shape"], case["seed"])
    fc, tc = seg.segment(vol, if_output_prob_map=True, if_output_itk=True)
    fcm, tcm = seg.segment(vol, if_output_prob_map=False, if_output_itk=False)
    return vol, fc.arr, tc.arr, fcm, tcm


def main():
    install_itk_stub()
    torch.set_num_threads(os.cpu_count())
    for name, case in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            vol, fc, tc, fcm, tcm = run_reference(case, tmp)
        assert fc.dtype == np.float64 and fc.shape == case["shape"]
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, fc=fc.astype(np.float32), tc=tc.astype(np.float32),
                            fc_mask=fcm.astype(np.uint8), tc_mask=tcm.astype(np.uint8),
                            meta=json.dumps({k: v for k, v in case.items()}))
        inner = fc[case["overlap"][2]:-case["overlap"][2], case["overlap"][0]:-case["overlap"][0],
                   case["overlap"][1]:-case["overlap"][1]]
        print(name, "fc mean %.4f std %.4f  mask frac %.3f  -> %s (%d KB)" % (
            inner.mean(), inner.std(), fcm.mean(), out, os.path.getsize(out) // 1024))


if __name__ == "__main__":
    main()
