"""One warp_volume launch at 384^3 for C = 1 and C = 8, and one two-field composition (for ncu --set full)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402
from scripts.bench_warp import smooth_disp  # noqa: E402

n = 384
eye = (np.eye(3), np.zeros(3))
disp = smooth_disp(n)
field = disp.permute(1, 2, 3, 0).flip(-1).contiguous()
for C in (1, 8):
    src = torch.rand(C, n, n, n, device="cuda")
    out = torch.empty_like(src)
    for _ in range(2):
        ops.warp_volume(src, field, eye, eye, (n, n, n), out=out)
    del src, out
u = [(disp / (n - 1)).contiguous(), (smooth_disp(n, 2.0, 8) / (n - 1)).contiguous()]
for _ in range(2):
    ops.compose((n, n, n), u, False)
torch.cuda.synchronize()
