#!/usr/bin/env python
"""Benchmark of the per-knee hot path (BASELINE.json metric: knee volumes/s for seg + ICON reg + warps).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo's CUDA path (one JSON line)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1  # the reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                             # N ranks, knees sharded by volume, no collective

One "step" = one synthetic 160x384x384 knee through BASELINE config 3: 3-D UNet segmentation (160 overlapping
32x128x128 tiles), GradICON registration to the atlas (both directions), warp of the FC/TC probability maps onto the
atlas grid and warp of 85 370 thickness-mesh vertices into atlas space.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATCH, OVERLAP = [128, 128, 32], (16, 16, 8)
SEG_FLOP_PER_VOLUME = 156.31e12      # 160 tiles x 976.94 GFLOP (every MAC the reference executes)
METRIC = "knee volumes/sec (seg+ICON reg+warp)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), source="measured (sustained)")
    except Exception:  # noqa: BLE001
        return dict(tflops=1400.0, hbm=6650.0, source="fallback")


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(suffix=".clocks.csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [v.strip() for v in line.split(",")]
                if len(c) < 8:
                    continue
                sm.append(float(c[1]))
                out["sm_max_mhz"] = float(c[2])
                for n, v in zip(names, c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)]
            out["sm_mhz"] = float(np.median(busy))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------- this repo's arm
def build_pipeline(device, tiles_per_batch=None):
    import torch

    from oai_analysis_2_b200 import synthetic
    from oai_analysis_2_b200.icon_registration import pretrained_models
    from oai_analysis_2_b200.pipeline import KneePipeline
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oai_analysis_2_b200.transforms import Geometry

    tmp = tempfile.mkdtemp(prefix="oai_bench_")
    cfg_json = os.path.join(tmp, "segmentation_train_config.pth.tar")
    with open(cfg_json, "w") as f:
        json.dump({"patch_size": PATCH, "model": "UNet",
                   "model_setting": {"in_channels": 1, "n_classes": 2, "bias": True, "BN": True}}, f)
    # ckpoint_path=None -> the reference's own random init (segmentation/utils.py:42-44 -> UNet.weights_init)
    seg_cfg = dict(ckpoint_path=None, training_config_file=cfg_json, device=str(device), batch_size=4,
                   overlap_size=OVERLAP, output_prob=True, output_itk=True, tiles_per_batch=tiles_per_batch)
    torch.manual_seed(1234)
    seg = Segmenter3DInPatchClassWise(mode="pred", config=seg_cfg)
    seg.pred_setup()
    torch.manual_seed(4321)
    reg = pretrained_models.OAI_knees_gradICON_model(pretrained=False)
    for net in reg.nets.values():  # icon zero-initialises lastConv (zero displacement); give the warps real work
        net._sd["lastConv.weight"].normal_(0, 0.02)
        net._sd["lastConv.bias"].normal_(0, 0.05)
        net._packed = None
    geom = Geometry(synthetic.OAI_SHAPE[::-1], synthetic.OAI_SPACING)
    atlas = synthetic.synthetic_knee(synthetic.OAI_SHAPE, seed=1)
    pipe = KneePipeline(seg, reg, atlas, geom, device)
    return pipe, geom


def make_inputs(rank, n_distinct=2):
    from oai_analysis_2_b200 import synthetic
    base = synthetic.synthetic_knee(synthetic.OAI_SHAPE, seed=100 + rank)
    vols = [base]
    for i in range(1, n_distinct):  # cheap distinct variants: mirrored / rolled copies of the blob field
        vols.append(np.ascontiguousarray(np.roll(base[:, ::-1, :], 17 * i, axis=2)))
    verts = np.concatenate([synthetic.synthetic_vertices(synthetic.N_VERTS_FC, seed=rank),
                            synthetic.synthetic_vertices(synthetic.N_VERTS_TC, seed=1000 + rank)])
    return vols, verts


def run_b200(args):
    import torch

    from oai_analysis_2_b200 import _lib, sharding

    rank, world, local = sharding.init_process_group()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pipe, geom = build_pipeline(device, args.tiles_per_batch)
    vols_h, verts_h = make_inputs(rank)
    vols_d = [torch.from_numpy(v).to(device) for v in vols_h]
    verts_d = torch.from_numpy(verts_h).to(device)
    vols_pin = [torch.from_numpy(v).pin_memory() for v in vols_h]
    verts_pin = torch.from_numpy(verts_h).pin_memory()

    import ctypes
    launches_per_step = None
    if not args.no_graph:
        # the whole per-knee path is one CUDA graph; the conv profiling events are recorded inside it
        mark = {}

        def on_record():
            _lib.lib.oai_profile_begin()
            mark["n0"] = _lib.launch_count()

        pipe.capture(vols_d[0].shape, geom, verts_d.shape[0], on_record)
        launches_per_step = _lib.launch_count() - mark["n0"]   # kernels recorded into the graph = launches per replay

    def step_device(i):
        if args.no_graph:
            return pipe.run_device(vols_d[i % len(vols_d)], geom, verts_d)
        return pipe.run_device_graph(vols_d[i % len(vols_d)], geom, verts_d)

    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sharding.barrier(world)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    if args.no_graph:
        _lib.lib.oai_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    torch.cuda.synchronize()
    conv_ms, conv_n, conv_fl, conv_xfl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib.oai_profile_end(ctypes.byref(conv_ms), ctypes.byref(conv_n), ctypes.byref(conv_fl),
                                        ctypes.byref(conv_xfl)), "profile")
    if args.no_graph:
        launches = _lib.launch_count() - n0
        prof_steps = args.steps
    else:  # the events inside the graph hold the timings of the last replay: one step's worth of conv launches
        launches = launches_per_step * args.steps
        prof_steps = 1
    sharding.barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), world)
    ms_per_step = ms / args.steps
    value = world * args.steps / (ms * 1e-3)

    # end to end through the public host API: pinned host volume in, atlas-space maps / fields / vertices out
    pipe.run(vols_pin[0], geom, verts_pin)  # allocates the pinned result buffers
    if not args.no_graph:
        for _ in pipe.run_stream((vols_pin[i % len(vols_pin)], verts_pin) for i in range(2)):
            pass                            # allocates the staging / double-buffered pinned buffers of the stream API
    sharding.barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.no_graph:
        for i in range(args.steps):
            res = pipe.run(vols_pin[i % len(vols_pin)], geom, verts_pin)
    else:
        # the public throughput API: every knee's pinned-host volume goes H2D and its results come back D2H inside the
        # timed region; the copies of neighbouring knees overlap the compute of the current one
        for res in pipe.run_stream(((vols_pin[i % len(vols_pin)], verts_pin) for i in range(args.steps))):
            checksum = float(res["vertices_atlas"][0, 0]) + float(res["FC_atlas"][80, 192, 192])  # touch the results
    torch.cuda.synchronize()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, world)
    sharding.barrier(world)
    e2e = dict(value=world * args.steps / e2e_s, unit="volumes/s", h2d_bytes_per_step=int(res["h2d_bytes"]),
               d2h_bytes_per_step=int(res["d2h_bytes"]), ms_per_step=1e3 * e2e_s / args.steps)

    peaks = load_peaks()
    traffic, traffic_src = None, None
    try:   # per-launch DRAM bytes of the conv kernel from the committed ncu capture of this same command
        with open(os.path.join(ROOT, "profiles", "r01_conv_dram_traffic.json")) as f:
            t = json.load(f)
        traffic, traffic_src = float(t["dram_bytes_per_launch"]), "profiles/r01_conv_dram_traffic.json"
    except Exception:  # noqa: BLE001
        pass
    conv_tflops = conv_fl.value / (conv_ms.value * 1e-3) / 1e12 if conv_ms.value > 0 else 0.0
    exec_tflops = conv_xfl.value / (conv_ms.value * 1e-3) / 1e12 if conv_ms.value > 0 else 0.0
    roofline = dict(bound="tensor", kernel="conv_igemm_kernel (tcgen05 implicit-GEMM conv3d)",
                    achieved=conv_tflops, peak=peaks["tflops"], unit="TFLOP/s", frac=conv_tflops / peaks["tflops"],
                    peak_source=peaks["source"] + " cuBLAS bf16 (fp16 runs at the same tcgen05 kind::f16 rate)",
                    traffic=traffic, traffic_source=traffic_src, launches_per_step=conv_n.value / prof_steps,
                    kernel_ms_per_step=conv_ms.value / prof_steps,
                    share_of_step=(conv_ms.value / prof_steps) / ms_per_step if ms else None,
                    algorithmic_flops_per_step=conv_fl.value / prof_steps,
                    executed_flops_per_step=conv_xfl.value / prof_steps, executed_tflops=exec_tflops,
                    executed_frac=exec_tflops / peaks["tflops"],
                    note="achieved = algorithmic FLOPs (every MAC the reference executes on its tile grid) / kernel "
                         "time; executed_* counts only the MACs issued after dead-halo elimination (decoder outputs "
                         "the kept tile interior does not depend on are skipped), i.e. the tensor-pipe utilisation")
    line = dict(metric=METRIC, value=value, unit="volumes/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="fp16 operands / fp32 accumulate (tcgen05 kind::f16); fp32 registration; fp64 warp coordinates",
                data="synthetic",
                config=dict(workload="BASELINE configs[2]: full per-knee path on one B200 (3-D UNet segmentation of a "
                                     "160x384x384 knee in 160 tiles of 32x128x128 + GradICON registration 80x192x192 "
                                     "both directions + FC/TC warp to the atlas grid + 85370-vertex warp); one knee per "
                                     "step per GPU, knees sharded by volume",
                            weights="random init (reference UNet.weights_init / icon default init)",
                            l2="per-step activation working set (tens of GB) >> 126 MB L2; inputs rotate over "
                               "distinct volumes",
                            seg_tflops_per_volume=SEG_FLOP_PER_VOLUME / 1e12),
                e2e=e2e, gpu_launches=int(launches * world), roofline=roofline, clocks=clocks)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_sample(vols_h[0], 1, 0)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sample(vol, n_seg_tiles, seed):
    """Time the oracle (CPU restatement of the reference path) on a bounded sample and extrapolate to one knee."""
    import torch

    from oracle import reg_oracle, seg_oracle, warp_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # -- segmentation: n_seg_tiles of the 160 tiles through the fp32 UNet
    sd = seg_oracle.make_unet_state_dict(1234, 1, 2, True, True, False)
    tiles, g = seg_oracle.partition(vol, PATCH, OVERLAP)
    n_tiles = tiles.shape[0]
    with torch.no_grad():
        t0 = time.perf_counter()
        torch.sigmoid(seg_oracle.unet_forward(sd, tiles[:n_seg_tiles], True))
        t_seg_tile = (time.perf_counter() - t0) / n_seg_tiles
    # -- registration: one direction of the lean cascade (the reference also evaluates discarded loss terms)
    rsd = reg_oracle.make_gradicon_state_dict(4321)
    nets = reg_oracle.split_state_dict(rsd)
    A = reg_oracle.resize_to_network(vol)
    B = reg_oracle.resize_to_network(vol[:, ::-1, :].copy())
    with torch.no_grad():
        t0 = time.perf_counter()
        phi = reg_oracle.final_map(reg_oracle.regis_net_forward(nets, A, B), reg_oracle.INPUT_SHAPE)
        t_reg_dir = time.perf_counter() - t0
    # -- ITK-style warp of the probability maps: a 2^20-voxel sample of one map
    geom = warp_oracle.Geometry(vol.shape[::-1], (0.3646, 0.3646, 0.7))
    tr = warp_oracle.CompositeTransform(reg_oracle.displacement_field_xyz(phi), geom, geom)
    nvox = int(np.prod(vol.shape))
    sample = 1 << 20
    W, H = vol.shape[2], vol.shape[1]
    lin = np.arange(sample) * (nvox // sample)
    j = np.stack([lin % W, (lin // W) % H, lin // (W * H)], -1)
    t0 = time.perf_counter()
    q = tr.transform_points(geom.index_to_physical(j))
    idx = geom.physical_to_index(q)
    warp_oracle._trilinear_clamped(vol.astype(np.float64), idx[..., ::-1])
    t_warp = (time.perf_counter() - t0) / sample
    t0 = time.perf_counter()
    tr.transform_points(geom.index_to_physical(j[:85370]))
    t_pts = time.perf_counter() - t0
    t_volume = t_seg_tile * n_tiles + 2 * t_reg_dir + 2 * nvox * t_warp + t_pts
    return dict(value=1.0 / t_volume, unit="volumes/s", cores=cores, kind="port",
                seconds_per_volume=t_volume,
                sample=f"oracle (torch fp32 / numpy f64 port of the reference path) on {cores} host threads: "
                       f"{n_seg_tiles} of {n_tiles} UNet tiles ({t_seg_tile:.2f} s/tile), one of two GradICON "
                       f"directions ({t_reg_dir:.1f} s), 2^20 of {nvox} warp voxels, 85370 vertices; extrapolated "
                       f"to one knee")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oai_analysis_2_b200 import synthetic
    vol = synthetic.synthetic_knee(synthetic.OAI_SHAPE, seed=100)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(vol, 1, i)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    cb = dict(vals[-1], value=v)
    line = dict(impl="reference", metric=METRIC, value=v, unit="volumes/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / v, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="fp32 (torch CPU) / fp64 (ITK-style warps)", data="synthetic",
                config=dict(workload="BASELINE configs[2] (same as the b200 arm); each step times a bounded sample of "
                                     "one knee and extrapolates", note="icon_registration / itk are not installable "
                                     "offline, so the registration and warp legs run the oracle port; the UNet leg is "
                                     "the reference's own torch ops"),
                cpu_baseline=cb, e2e=dict(value=v, unit="volumes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--tiles-per-batch", type=int, default=None)
    ap.add_argument("--precision", default=None, choices=["fp16", "mixed", "fp16x2", "fp16x3", "bf16"],
                    help="segmentation precision plan (default: the product default, 'mixed')")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying the CUDA graph")
    args = ap.parse_args()
    if args.precision:
        os.environ["OAI_B200_SEG_PRECISION"] = args.precision
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
