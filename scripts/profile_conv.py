"""Run one representative segmentation conv layer a few times (for ncu captures and per-layer timing).

    python scripts/profile_conv.py dc2 [NT]       # layers: ec1 dc2 dc1 dc5 dc4 dc8 ec3 ec5 ec7 dc3
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402

LAYERS = {  # name: (D, H, W, c0, c1, cout, pointwise, region)
    "ec1": (32, 128, 128, 32, 0, 64, False, None),
    "dc2": (32, 128, 128, 128, 64, 64, False, (7, 18, 15, 98)),
    "dc2full": (32, 128, 128, 128, 64, 64, False, None),
    "dc1": (32, 128, 128, 64, 0, 64, False, (8, 16, 16, 96)),
    "dc1full": (32, 128, 128, 64, 0, 64, False, None),
    "ec2": (16, 64, 64, 64, 0, 64, False, None),
    "ec3": (16, 64, 64, 64, 0, 128, False, None),
    "dc5": (16, 64, 64, 256, 128, 128, False, (2, 12, 6, 52)),
    "dc5full": (16, 64, 64, 256, 128, 128, False, None),
    "dc4": (16, 64, 64, 128, 0, 128, False, (3, 10, 7, 50)),
    "ec4": (8, 32, 32, 128, 0, 128, False, None),
    "ec5": (8, 32, 32, 128, 0, 256, False, None),
    "dc8": (8, 32, 32, 512, 256, 256, False, None),
    "dc7": (8, 32, 32, 256, 0, 256, False, None),
    "ec6": (4, 16, 16, 256, 0, 256, False, None),
    "ec7": (4, 16, 16, 256, 0, 512, False, None),
    "dc3": (16, 64, 64, 128, 0, 128, True, (3, 10, 7, 50)),
    "dc9": (4, 16, 16, 512, 0, 512, True, None),
}


def run(name, NT, iters=5, flags=0):
    D, H, W, c0, c1, cout, pw, region = LAYERS[name]
    dev = "cuda"
    x0 = torch.randn(NT, D, H, W, c0, device=dev).half()
    x1 = torch.randn(NT, D, H, W, c1, device=dev).half() if c1 else None
    cin = c0 + c1
    w = torch.randn(*((cout, cin) if pw else (cout, cin, 3, 3, 3))) * (1.0 / (cin * (1 if pw else 27)) ** 0.5)
    wp = ops.pack_conv_weights(w, c0, c1, D, H, W, pw, 0)
    bias = torch.zeros(cout, device=dev)
    out = torch.empty((NT, D, H, W, cout), dtype=torch.float16, device=dev)
    for _ in range(2):
        ops.conv3d_igemm(x0, x1, wp, bias, cout, pw, True, 0, out=out, region=region, flags=flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv3d_igemm(x0, x1, wp, bias, cout, pw, True, 0, out=out, region=region, flags=flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    frac = 1.0 if region is None else region[1] * region[3] / (D * H)
    fl = 2.0 * NT * D * H * W * cout * cin * (1 if pw else 27)
    return dict(layer=name, NT=NT, flags=flags, ms=ms, algorithmic_tflops=fl / ms / 1e9, executed_tflops=fl * frac / ms / 1e9,
                plan=ops.conv_plan(D if region is None else region[1], H, W, c0, c1, cout, pw))


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(LAYERS)
    NT = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    flag_list = [int(f) for f in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    for n in names:
        for f in flag_list:
            for rep in range(2):
                r = run(n, NT, flags=f)
            print(json.dumps(r)[:150], flush=True)
