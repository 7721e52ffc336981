"""CPU tier: UNet.weights_init (reference oai_analysis/segmentation/networks.py:71-78, reached through
initialize_model(ckpoint_path=None), utils.py:42-44): xavier_normal_ on every conv weight, zero bias, BatchNorm left at
its defaults -- checked statistically, and draw for draw against the reference's own module when /root/reference is
present (build container only)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oai_analysis_2_b200.segmentation.networks import UNet
from oai_analysis_2_b200.segmentation.utils import initialize_model


def test_weights_init_statistics():
    torch.manual_seed(7)
    net = UNet(1, 2, bias=True, BN=True)
    assert all(float(v.abs().max()) == 0 for k, v in net.state_dict().items() if k.endswith(".0.weight"))
    initialize_model(net, ckpoint_path=None)        # -> weights_init()
    sd = net.state_dict()
    for key, w in sd.items():
        if key.endswith(".0.weight") or key == "dc0.weight":
            k3 = int(np.prod(w.shape[2:]))
            # torch's fan computation on the STORED shape: fan_in = size(1) k^3, fan_out = size(0) k^3 (for a
            # ConvTranspose3d weight [cin, cout, ...] that swaps the roles, exactly as in the reference)
            std = math.sqrt(2.0 / ((w.shape[0] + w.shape[1]) * k3))
            n = w.numel()
            assert abs(float(w.std()) - std) < 5 * std / math.sqrt(2 * n) + 1e-3 * std, key
            assert abs(float(w.mean())) < 5 * std / math.sqrt(n), key
        elif key.endswith(".0.bias") or key == "dc0.bias":
            assert float(w.abs().max()) == 0, key
        elif key.endswith(".1.weight") or key.endswith(".1.running_var"):
            assert torch.equal(w, torch.ones_like(w)), key
        elif key.endswith(".1.bias") or key.endswith(".1.running_mean"):
            assert torch.equal(w, torch.zeros_like(w)), key


def test_weights_init_invalidates_packed_handles_and_keeps_keys():
    net = UNet(1, 3, bias=False, BN=False)
    keys = set(net.state_dict())
    assert "ec0.0.bias" not in keys and "ec0.1.weight" not in keys and net.state_dict()["dc0.weight"].shape == (3, 64, 1, 1, 1)
    net._handles["stale"] = object()
    net.weights_init()
    assert net._handles == {} and set(net.state_dict()) == keys


@pytest.mark.skipif(not os.path.isdir("/root/reference/oai_analysis"), reason="reference sources not on this box")
def test_weights_init_draws_match_the_reference_module():
    """Same seed, same draw order: the state dict equals the reference UNet's after its own weights_init()."""
    sys.path.insert(0, "/root/reference")
    try:
        from oai_analysis.segmentation.networks import UNet as RefUNet
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(1234)
    ref = RefUNet(1, 2, bias=True, BN=True)
    torch.manual_seed(99)
    ref.weights_init()
    torch.manual_seed(99)
    net = UNet(1, 2, bias=True, BN=True)
    net.weights_init()
    rsd, sd = ref.state_dict(), net.state_dict()
    assert set(rsd) == set(sd)
    for k in rsd:
        if rsd[k].is_floating_point() and not (k.endswith(".0.bias") or k == "dc0.bias"):
            assert torch.equal(rsd[k], sd[k]), k
        elif k.endswith(".0.bias") or k == "dc0.bias":
            assert float(sd[k].abs().max()) == 0 and float(rsd[k].abs().max()) == 0
