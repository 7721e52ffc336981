"""Time single tallUNet2 up-path layers (oai_reg_convt4 / oai_reg_convt4_mma / oai_reg_convt4_umma) at the GradICON
shapes (half-resolution nets: 80x192x192 output; quarter-resolution nets: 40x96x96)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402

SHAPES = [(48, 16, (40, 96, 96)), (96, 32, (20, 48, 48)), (192, 64, (10, 24, 24)), (512, 128, (5, 12, 12)),
          (512, 256, (3, 6, 6)), (48, 16, (20, 48, 48)), (96, 32, (10, 24, 24)), (192, 64, (5, 12, 12))]
which = [int(v) for v in sys.argv[1:]] or range(len(SHAPES))
for i in which:
    cin, cout, dims = SHAPES[i]
    N = 2
    x = torch.randn(N, cin, *dims, device="cuda")
    w = (torch.randn(cin, 64, cout, device="cuda") * 0.05).contiguous()
    b, s, t = torch.randn(cout, device="cuda"), torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    out = torch.empty(N, cout, *[2 * d for d in dims], device="cuda")
    wpk, wexp = ops.reg_pack_convt4(w, cin, cout)
    res = {}
    variants = [("mma", lambda: ops.reg_convt4(x, cin, w, b, s, t, out, cout, wpk=wpk, wexp=wexp)),
                ("fp32", lambda: ops.reg_convt4(x, cin, w, b, s, t, out, cout))]
    if cout <= 128 and dims[2] >= 8:
        wu = ops.reg_pack_convt4_umma(w, cin, cout, wexp)
        variants.insert(0, ("umma", lambda: ops.reg_convt4_umma(x, cin, wu, wexp, b, s, t, out, cout)))
    if os.environ.get("OAI_BENCH_ONLY"):
        variants = [v for v in variants if v[0] in os.environ["OAI_BENCH_ONLY"].split(",")]
    for name, fn in variants:
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res["ms_" + name] = e0.elapsed_time(e1) / 10
    gmac = N * cin * cout * 64 * dims[0] * dims[1] * dims[2] / 1e9
    rec = dict(cin=cin, cout=cout, dims=dims, gmac=gmac, debug=os.environ.get("OAI_CONVT4_DEBUG", "0"), **res)
    for k in list(res):
        rec["tflops_" + k[3:]] = 2 * gmac / res[k]
    print(json.dumps(rec), flush=True)
