"""Run the GradICON registration cascade a few times on synthetic 160x384x384 volumes (for ncu / timing)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import synthetic  # noqa: E402
from oai_analysis_2_b200.icon_registration import itk_wrapper, pretrained_models  # noqa: E402

torch.manual_seed(4321)
model = pretrained_models.OAI_knees_gradICON_model(pretrained=False)
for net in model.nets.values():
    net._sd["lastConv.weight"].normal_(0, 0.02)
    net._sd["lastConv.bias"].normal_(0, 0.05)
    net._packed = None
A = torch.from_numpy(synthetic.synthetic_knee((160, 384, 384), 3, n_blobs=16)).cuda()
B = A.flip(1).contiguous()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    itk_wrapper.register_pair_device(model, A, B)
    torch.cuda.synchronize()
    print(f"register_pair_device: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
