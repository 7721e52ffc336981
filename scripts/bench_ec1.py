"""ec1-shaped layer (32 -> 64 at 32x128x128, 160 tiles) under the planner's A/B flags and term counts."""
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
    NT, D, H, W = 160, 32, 128, 128
    for c0, cout in ((32, 64), (64, 64)):
        x = torch.randn(NT, D, H, W, c0, device="cuda").half()
        xs = torch.cat((x, torch.zeros_like(x)), -1).contiguous()
        w = torch.randn(cout, c0, 3, 3, 3) * 0.05
        b = torch.zeros(cout, device="cuda")
        for name, flags, terms, split_out in (("default", 0, 1, False), ("one kh row per stage", 64, 1, False), ("wide rows (128 B)", 32, 1, False),
                                              ("per-tap", 1, 1, False), ("terms 2", 0, 2, False),
                                              ("default, split out", 0, 1, True)):
            if flags in (32, 64) and c0 != 32:
                continue
            wp = ops.pack_conv_weights_ex(w, c0, 0, D, H, W, 0, terms, 0, flags)
            src = xs if terms > 1 else x
            ms = timeit(lambda: ops.conv3d_igemm_ex(src, None, wp, b, cout, c0, 0, 0, True, 0, terms,
                                                    in_split=terms > 1, out_split=split_out, flags=flags))
            fl = 2.0 * NT * D * H * W * cout * c0 * 27 * (2 if terms == 2 else 1)
            print(json.dumps(dict(cin=c0, cout=cout, variant=name, ms=round(ms, 3),
                                  issued_tflops=round(fl / ms / 1e9, 1))), flush=True)
        del x, xs


if __name__ == "__main__":
    main()
