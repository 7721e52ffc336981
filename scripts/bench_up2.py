"""k2s2 transposed-conv layers (dc6 / dc3 / dc9 shapes, 160 full tiles) under the planner's A/B flag 128 = no accumulator
ping-pong / stationary weights.  (A third variant, the epilogue with its global stores compiled out, was measured once:
profiles/r02_bench_up2_variants.jsonl.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oai_analysis_2_b200 import ops  # noqa: E402


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    NT = 160
    for name, cin, cout, dims in (("dc6", 256, 256, (8, 32, 32)), ("dc3", 128, 128, (16, 64, 64)),
                                  ("dc9", 512, 512, (4, 16, 16))):
        D, H, W = dims
        x = torch.randn(NT, D, H, W, cin, device="cuda").half()
        w = torch.randn(cout, cin, 2, 2, 2) * 0.05
        b = torch.zeros(cout, device="cuda")
        for variant, flags in (("default", 0), ("no ping-pong", 128)):
            wp = ops.pack_conv_weights_ex(w, cin, 0, D, H, W, 2, 1, 0, flags)
            ms = timeit(lambda: ops.conv3d_igemm_ex(x, None, wp, b, cout, cin, 0, 2, True, 0, 1, flags=flags))
            fl = 2.0 * NT * D * H * W * cout * cin * 8
            print(json.dumps(dict(layer=name, variant=variant, ms=round(ms, 3), issued_tflops=round(fl / ms / 1e9, 1),
                                  out_gb=round(NT * 8 * D * H * W * cout * 2 / 1e9, 2))), flush=True)
        del x


if __name__ == "__main__":
    main()
