"""CPU tier: the weight packer + the kernel's issue-loop arithmetic (software model) against torch conv3d."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conv_emulator import emulate
from oai_analysis_2_b200 import ops


def _case(NT, D, H, W, c0, c1, cout, pointwise, flags=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(NT, D, H, W, c0, generator=g).half()
    x1 = torch.randn(NT, D, H, W, c1, generator=g).half() if c1 else None
    cin = c0 + c1
    if pointwise:
        w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).half().float()
    else:
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5).half().float()
    bias = torch.randn(cout, generator=g)
    plan = ops.conv_plan(D, H, W, c0, c1, cout, pointwise, flags)
    wpack = ops.pack_conv_weights(w, c0, c1, D, H, W, pointwise, 0, flags, device="cpu").numpy()
    k16 = 4 if (c1 or c0 > 32) else (2 if c0 > 16 else 1)
    got = emulate(x0.numpy(), None if x1 is None else x1.numpy(), wpack, bias.numpy(), plan, cout, True, k16)
    x = x0 if x1 is None else torch.cat((x0, x1), -1)
    xn = x.float().permute(0, 4, 1, 2, 3)
    wn = w.view(cout, cin, 1, 1, 1) if pointwise else w
    ref = F.relu(F.conv3d(xn, wn, bias, padding=0 if pointwise else 1)).permute(0, 2, 3, 4, 1).numpy()
    return got, ref, plan


@pytest.mark.parametrize("args", [
    (1, 4, 2, 128, 64, 0, 64, False, 0),     # row-shared, R=4? (D=4)
    (1, 8, 2, 128, 32, 0, 64, False, 0),     # row-shared, 32-channel source (k16=2), R=8
    (1, 2, 1, 128, 64, 64, 64, False, 0),    # row-shared, two sources
    (1, 4, 2, 128, 64, 0, 64, False, 1),     # same geometry forced per-tap
    (2, 4, 4, 64, 64, 0, 128, False, 0),     # per-tap, kd stacked (N=256+128), R=4
    (1, 4, 8, 32, 64, 64, 128, False, 0),    # per-tap two sources
    (1, 2, 8, 16, 64, 0, 256, False, 0),     # per-tap kd_per_block=1, R=2
    (1, 2, 8, 16, 64, 0, 512, False, 0),     # two N halves
    (1, 4, 4, 32, 128, 0, 128, True, 0),     # pointwise
    (1, 8, 2, 64, 64, 0, 64, False, 0),      # per-tap cout 64: R=8, N=192 stacks
])
def test_emulated_kernel_matches_conv3d(args):
    got, ref, plan = _case(*args)
    err = np.abs(got - ref).max()
    assert err < 2e-3, (err, plan)
