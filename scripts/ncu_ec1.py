"""One ec1-shaped launch (32 -> 64 at 32x128x128) for an ncu --set full capture of conv_igemm_kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oai_analysis_2_b200 import ops  # noqa: E402

NT, D, H, W = 40, 32, 128, 128
c0, cout = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 64
x = torch.randn(NT, D, H, W, c0, device="cuda").half()
w = torch.randn(cout, c0, 3, 3, 3) * 0.05
b = torch.zeros(cout, device="cuda")
wp = ops.pack_conv_weights_ex(w, c0, 0, D, H, W, 0, 1)
for _ in range(3):
    ops.conv3d_igemm_ex(x, None, wp, b, cout, c0)
torch.cuda.synchronize()
