"""CPU tier: the segmentation oracle (oracle/seg_oracle.py) against fixtures produced by the reference's own code
(tests/golden/make_golden.py ran oai_analysis/segmentation/segmenter.py unmodified)."""
import numpy as np
import pytest
import torch

from helpers import dice, load_golden
from oracle.seg_oracle import make_unet_state_dict, partition, segment, synthetic_knee, tiling_geometry


@pytest.mark.parametrize("name", ["seg_small_pertap", "seg_small_nobn"])
def test_oracle_matches_reference_golden(name):
    z, m = load_golden(name)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = make_unet_state_dict(m["seed"], 1, 2, m["bias"], m["BN"], True, m["head_gain"], m["head_bias"])
    vol = synthetic_knee(tuple(m["shape"]), m["seed"])
    fc, tc = segment(vol, sd, m["patch"], tuple(m["overlap"]), 4, m["BN"])
    assert fc.dtype == np.float64 and fc.shape == tuple(m["shape"])
    # float32 fixtures of a float64 map: exact up to the fixture's own rounding / CPU summation order
    assert np.abs(fc - z["fc"]).max() < 2e-5
    assert np.abs(tc - z["tc"]).max() < 2e-5
    assert dice(fc, z["fc_mask"]) > 0.9999 and dice(tc, z["tc_mask"]) > 0.9999
    # the mask is a real structure, not all-0/all-1
    assert 0.2 < z["fc_mask"].sum() / (z["fc"] > 0).sum() < 0.8


def test_tiling_geometry_production_case():
    """160x384x384 with patch [128,128,32], overlap (16,16,8) -> 10x4x4 tiles, pads (8,8),(16,16),(16,16)."""
    g = tiling_geometry((160, 384, 384), [128, 128, 32], (16, 16, 8))
    assert g["tile"].tolist() == [32, 128, 128] and g["effective"].tolist() == [16, 96, 96]
    assert g["grid"].tolist() == [10, 4, 4] and g["padded"].tolist() == [16, 32, 32]


def test_partition_reflect_and_order():
    vol = np.arange(6 * 10 * 12, dtype=np.float32).reshape(6, 10, 12)
    tiles, g = partition(vol, [8, 8, 4], (2, 2, 1))
    assert tiles.shape == (int(np.prod(g["grid"])), 1, 4, 8, 8)
    padded = np.pad(vol, [(1, g["padded"][0] - 1), (2, g["padded"][1] - 2), (2, g["padded"][2] - 2)], mode="reflect")
    eff = g["effective"]
    idx = (1 * g["grid"][1] + 0) * g["grid"][2] + 2
    np.testing.assert_array_equal(tiles[idx, 0].numpy(), padded[eff[0]:eff[0] + 4, 0:8, 2 * eff[2]:2 * eff[2] + 8])
