"""Launch the warp / composition kernels once each at benchmark shapes (for ncu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402
from scripts.bench_warp import smooth_disp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
eye = (np.eye(3), np.zeros(3))
disp = smooth_disp(n)
field = disp.permute(1, 2, 3, 0).flip(-1).contiguous()
src = torch.rand(2, n, n, n, device="cuda")
for _ in range(2):
    ops.warp_volume(src, field, eye, eye, (n, n, n))
u = [(disp / (n - 1)).contiguous(), (smooth_disp(n, 2.0, 8) / (n - 1)).contiguous()]
for _ in range(2):
    ops.compose((n, n, n), u, False)
# pipeline shape: 160x384x384 output through an 80x192x192 field (scale 0.5 affine)
f2 = smooth_disp(192)[:, :80].permute(1, 2, 3, 0).flip(-1).contiguous()
half = (np.diag([0.5, 0.5, 0.5]), np.array([-0.25, -0.25, -0.25]))
two = (np.diag([2.0, 2.0, 2.0]), np.array([0.5, 0.5, 0.5]))
src2 = torch.rand(2, 160, 384, 384, device="cuda")
for _ in range(2):
    ops.warp_volume(src2, f2, half, two, (160, 384, 384))
torch.cuda.synchronize()
