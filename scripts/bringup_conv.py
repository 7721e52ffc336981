"""GPU bring-up of the tcgen05 conv kernel: each case runs in its own process (a trap poisons the CUDA context)
under a timeout, compares against torch conv3d (fp32, TF32 off) and prints one line.

    python scripts/bringup_conv.py            # all cases
    python scripts/bringup_conv.py --case 3   # one case, in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, NT, D, H, W, c0, c1, cout, pointwise, flags, ab_format)
CASES = [
    ("pointwise_64_64_w16", 2, 4, 8, 16, 64, 0, 64, True, 0, 0),
    ("pointwise_128_128_w32", 2, 4, 4, 32, 128, 0, 128, True, 0, 0),
    ("pertap_kd1_64_256_w16", 2, 4, 16, 16, 64, 0, 256, False, 0, 0),
    ("pertap_kd1_64_512_w16", 2, 4, 16, 16, 64, 0, 512, False, 0, 0),
    ("pertap_kd3_128_128_w32", 2, 8, 32, 32, 128, 0, 128, False, 0, 0),
    ("pertap_kd3_64_64_w64", 2, 16, 64, 64, 64, 0, 64, False, 0, 0),
    ("pertap_kd3_dual_256_128_128_w64", 1, 16, 64, 64, 256, 128, 128, False, 0, 0),
    ("pertap_forced_64_64_w128", 2, 8, 4, 128, 64, 0, 64, False, 1, 0),
    ("rowshared_64_64_w128", 2, 8, 4, 128, 64, 0, 64, False, 0, 0),
    ("rowshared_c32_64_w128", 2, 8, 4, 128, 32, 0, 64, False, 0, 0),
    ("rowshared_dual_128_64_64_w128", 1, 32, 8, 128, 128, 64, 64, False, 0, 0),
    ("rowshared_bf16", 2, 8, 4, 128, 64, 0, 64, False, 0, 1),
    ("pertap_big_tile_64_64", 4, 32, 128, 128, 64, 0, 64, False, 1, 0),
    ("rowshared_big_tile_64_64", 4, 32, 128, 128, 64, 0, 64, False, 0, 0),
    ("up2_128_128_w64", 2, 16, 64, 64, 128, 0, 128, 2, 0, 0),
    ("up2_256_256_w32", 2, 8, 32, 32, 256, 0, 256, 2, 0, 0),
    ("up2_512_512_w16", 2, 4, 16, 16, 512, 0, 512, 2, 0, 0),
    ("up2_64_64_w16", 2, 4, 8, 16, 64, 0, 64, 2, 0, 0),
]


def run_case(i):
    import torch
    import torch.nn.functional as F

    from oai_analysis_2_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, NT, D, H, W, c0, c1, cout, pw, flags, fmt = CASES[i]
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(i)
    dev = "cuda"
    x0 = torch.randn(NT, D, H, W, c0, generator=g, device=dev).to(dt)
    x1 = torch.randn(NT, D, H, W, c1, generator=g, device=dev).to(dt) if c1 else None
    cin = c0 + c1
    if pw == 2:
        w = (torch.randn(cout, cin, 2, 2, 2, generator=g, device=dev) / cin ** 0.5).to(dt).float()
        bias = torch.randn(cout, generator=g, device=dev)
        wpack = ops.pack_convt2_weights(w, D, H, W, fmt)
        out = ops.convt2_igemm(x0, wpack, bias, cout, True, fmt)
        torch.cuda.synchronize()
        ref = F.relu(F.conv_transpose3d(x0.float().permute(0, 4, 1, 2, 3), w.transpose(0, 1).contiguous(), bias,
                                        stride=2)).permute(0, 2, 3, 4, 1)
        err = (out.float() - ref).abs()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.convt2_igemm(x0, wpack, bias, cout, True, fmt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("RESULT " + json.dumps(dict(case=name, ok=bool(err.max().item() < 2e-2), max_err=err.max().item(),
                                          ms=ms, tflops=2.0 * NT * D * H * W * 8 * cout * cin / ms / 1e9,
                                          plan=ops.conv_plan(D, H, W, cin, 0, cout, 2))), flush=True)
        return
    if pw:
        w = (torch.randn(cout, cin, generator=g, device=dev) / cin ** 0.5).to(dt).float()
    else:
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g, device=dev) / (27 * cin) ** 0.5).to(dt).float()
    bias = torch.randn(cout, generator=g, device=dev)
    plan = ops.conv_plan(D, H, W, c0, c1, cout, pw, flags)
    wpack = ops.pack_conv_weights(w, c0, c1, D, H, W, pw, fmt, flags)
    out = ops.conv3d_igemm(x0, x1, wpack, bias, cout, pw, True, fmt, flags=flags)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat((x0, x1), -1)
    xn = x.float().permute(0, 4, 1, 2, 3)
    wn = w.view(cout, cin, 1, 1, 1) if pw else w
    ref = F.relu(F.conv3d(xn, wn, bias, padding=0 if pw else 1)).permute(0, 2, 3, 4, 1)
    err = (out.float() - ref).abs()
    tol = 2e-2 if fmt == 0 else 6e-2
    # timing
    ms = None
    try:
        for _ in range(2):
            ops.conv3d_igemm(x0, x1, wpack, bias, cout, pw, True, fmt, out=out, flags=flags)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv3d_igemm(x0, x1, wpack, bias, cout, pw, True, fmt, out=out, flags=flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    except Exception as ex:  # noqa
        ms = str(ex)
    flops = 2.0 * NT * D * H * W * cout * cin * (1 if pw else 27)
    res = dict(case=name, ok=bool(err.max().item() < tol), max_err=err.max().item(), mean_err=err.mean().item(),
               ref_absmax=ref.abs().max().item(), plan=plan, ms=ms,
               tflops=(flops / (ms * 1e-3) / 1e12) if isinstance(ms, float) else None)
    if not res["ok"]:
        bad = (err > tol).nonzero()
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:6].tolist()
        res["bad_frac"] = bad.shape[0] / err.numel()
    print("RESULT " + json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--only", type=str, default=None, help="comma-separated case indices")
    ap.add_argument("--timeout", type=int, default=120)
    a = ap.parse_args()
    if a.case is not None:
        run_case(a.case)
        return
    idx = range(len(CASES)) if a.only is None else [int(s) for s in a.only.split(",") if int(s) < len(CASES)]
    for i in idx:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True,
                               text=True, timeout=a.timeout)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1], flush=True)
            else:
                tail = (r.stdout[-600:] + r.stderr[-1200:]).replace("\n", " | ")
                print(f"FAILED case {i} {CASES[i][0]} rc={r.returncode}: {tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT case {i} {CASES[i][0]} after {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    main()
