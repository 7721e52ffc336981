"""BASELINE configs[4]: trilinear warp / composition micro-benchmark sweep (HBM GB/s against the measured roofline).

    python scripts/bench_warp.py            # prints one JSON line per case

warp      : ITK-semantics resample (oai_warp_volume): per output voxel 12 B of field + 4C B of image read, 4C B written
compose   : c <- c + S(u, c) for two 3-channel fields (oai_compose): 36 B/voxel (2 x 12 B fields read, 12 B written)
Algorithmic bytes follow SURVEY §8(d): warp (12 + 8 C) N^3, composition 36 N^3.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oai_analysis_2_b200 import ops  # noqa: E402


def smooth_disp(n, sigma_vox=2.0, seed=7):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.randn(1, 3, max(2, n // 8), max(2, n // 8), max(2, n // 8), generator=g, device="cuda")
    d = torch.nn.functional.interpolate(lo, size=(n, n, n), mode="trilinear", align_corners=True)[0]
    return (d * sigma_vox).contiguous()


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = 6547.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2, rewritten between cases
    eye = (np.eye(3), np.zeros(3))
    sizes = [int(v) for v in sys.argv[1:]] or [64, 96, 128, 192, 256, 384]
    for n in sizes:
        disp = smooth_disp(n)                                   # [3,n,n,n] voxels
        field = disp.permute(1, 2, 3, 0).flip(-1).contiguous()  # [n,n,n,3] x,y,z components
        for C in (1, 2, 3, 4, 8):
            src = torch.rand(C, n, n, n, device="cuda")
            flush.zero_()
            ms = timeit(lambda: ops.warp_volume(src, field, eye, eye, (n, n, n)))
            b = (12 + 8 * C) * n ** 3
            print(json.dumps(dict(op="warp_volume", n=n, C=C, ms=ms, gbs=b / ms / 1e6, frac_of_measured_hbm=b / ms / 1e6 / peak)),
                  flush=True)
        u = [(disp / (n - 1)).contiguous(), (smooth_disp(n, 2.0, 8) / (n - 1)).contiguous()]
        flush.zero_()
        ms = timeit(lambda: ops.compose((n, n, n), u, False))
        b = 36 * n ** 3
        print(json.dumps(dict(op="compose2", n=n, ms=ms, gbs=b / ms / 1e6, frac_of_measured_hbm=b / ms / 1e6 / peak)), flush=True)
    # intensity windowing of one knee: 3 histogram reads + 1 apply read + 1 write = 20 B / voxel
    vol = torch.rand(160, 384, 384, device="cuda") * 900.0
    out = torch.empty_like(vol)
    flush.zero_()
    ms = timeit(lambda: ops.intensity_window(vol, 0.1, 99.9, 0.0, 1.0, out=out))
    b = 20 * vol.numel()
    print(json.dumps(dict(op="intensity_window", n=[160, 384, 384], ms=ms, gbs=b / ms / 1e6,
                          frac_of_measured_hbm=b / ms / 1e6 / peak)), flush=True)


if __name__ == "__main__":
    main()
