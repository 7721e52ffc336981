"""Time single tallUNet2 down-path layers (oai_reg_conv3 on the fp32 CUDA-core kernels / split-K GEMMs, oai_reg_conv3_umma
on tcgen05) at the GradICON shapes (half-resolution nets: 80x192x192 input; quarter-resolution nets: 40x96x96)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402

SHAPES = [(16, 32, (40, 96, 96)), (32, 64, (20, 48, 48)), (64, 256, (10, 24, 24)), (16, 32, (20, 48, 48)),
          (32, 64, (10, 24, 24))]
which = [int(v) for v in sys.argv[1:]] or range(len(SHAPES))
for i in which:
    cin, cout, dims = SHAPES[i]
    N = 2
    x = torch.randn(N, cin, *dims, device="cuda")
    w = (torch.randn(cin, 27, cout, device="cuda") * 0.05).contiguous()
    b = torch.randn(cout, device="cuda")
    out = torch.empty(N, cout, *[(d + 1) // 2 for d in dims], device="cuda")
    wu, wexp = ops.reg_pack_conv3_umma(w, cin, cout)
    res = {}
    for name, fn in (("umma", lambda: ops.reg_conv3_umma(x, cin, wu, wexp, b, out, cout)),
                     ("fp32", lambda: ops.reg_conv3(x, cin, w, b, out, cout, 2, True, True))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res["ms_" + name] = e0.elapsed_time(e1) / 10
    gmac = N * cin * cout * 27 * out.shape[2] * out.shape[3] * out.shape[4] / 1e9
    print(json.dumps(dict(cin=cin, cout=cout, dims=dims, gmac=gmac, **res)), flush=True)
