"""CPU tier: the dead-halo regions the B200 path skips (UNet.needed_regions) really are dead: poisoning every decoder
activation outside its region with NaN leaves the kept tile interior bit-identical."""
import numpy as np
import torch
import torch.nn.functional as F

from oai_analysis_2_b200.segmentation.networks import UNet
from oracle.seg_oracle import UNET_LAYERS, _block, make_unet_state_dict, unet_forward


def _poison(x, box):
    if box is None:
        return x
    lo, hi = box
    y = torch.full_like(x, float("nan"))
    sl = (slice(None), slice(None)) + tuple(slice(int(a), int(b) + 1) for a, b in zip(lo, hi))
    y[sl] = x[sl]
    return y


def test_interior_does_not_depend_on_skipped_regions():
    tile, overlap = (16, 64, 32), (4, 8, 8)
    sd = make_unet_state_dict(3, 1, 2, True, True, True)
    kinds = {n: k for n, k, _, _ in UNET_LAYERS}
    B = UNet.needed_regions(tile, overlap)
    x = torch.rand(1, 1, *tile)
    blk = lambda n, t: _block(sd, n, kinds[n], t, True)  # noqa: E731
    up = lambda n, t: _poison(blk(n, t), (B[n][0] * 2, B[n][1] * 2 + 1))  # noqa: E731  (region given on the input grid)
    with torch.no_grad():
        ref = unet_forward(sd, x, True)
        syn0 = blk("ec1", blk("ec0", x))
        syn1 = blk("ec3", blk("ec2", F.max_pool3d(syn0, 2)))
        syn2 = blk("ec5", blk("ec4", F.max_pool3d(syn1, 2)))
        e7 = blk("ec7", blk("ec6", F.max_pool3d(syn2, 2)))
        d8 = _poison(blk("dc8", torch.cat((up("dc9", e7), syn2), 1)), B["dc8"])
        d7 = _poison(blk("dc7", d8), B["dc7"])
        d5 = _poison(blk("dc5", torch.cat((up("dc6", d7), syn1), 1)), B["dc5"])
        d4 = _poison(blk("dc4", d5), B["dc4"])
        d2 = _poison(blk("dc2", torch.cat((up("dc3", d4), syn0), 1)), B["dc2"])
        d1 = _poison(blk("dc1", d2), B["dc1"])
        out = F.conv3d(d1, sd["dc0.weight"], sd["dc0.bias"])
    o = overlap
    inner = (slice(None), slice(None), slice(o[0], tile[0] - o[0]), slice(o[1], tile[1] - o[1]),
             slice(o[2], tile[2] - o[2]))
    assert not torch.isnan(out[inner]).any()
    assert torch.equal(out[inner], ref[inner])
    # the regions are genuinely smaller than the tile at the top levels
    frac = np.prod(B["dc2"][1] - B["dc2"][0] + 1) / np.prod(tile)
    assert frac < 0.6


def test_production_geometry_regions():
    B = UNet.needed_regions((32, 128, 128), (8, 16, 16))
    assert B["dc1"][0].tolist() == [8, 16, 16] and B["dc1"][1].tolist() == [23, 111, 111]
    assert B["dc2"][0].tolist() == [7, 15, 15] and B["dc3"][0].tolist() == [3, 7, 7] and B["dc3"][1].tolist() == [12, 56, 56]
    assert UNet._region_arg(B["dc2"], (32, 128, 128), 64) == (7, 18, 15, 98)
    assert UNet._region_arg(B["dc1"], (32, 128, 128), 64) == (8, 16, 16, 96)
    assert UNet._region_arg(B["dc4"], (16, 64, 64), 128) == (3, 10, 7, 50)
    assert UNet._region_arg(B["dc3"], (16, 64, 64), 128, True) == (3, 10, 7, 50)
