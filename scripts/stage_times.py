"""Per-stage device times of one knee through the hot path (CUDA events around each stage, after warm-up).

    python scripts/stage_times.py [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oai_analysis_2_b200 import ops  # noqa: E402
from oai_analysis_2_b200.icon_registration import itk_wrapper  # noqa: E402
from oai_analysis_2_b200.transforms import CompositeTransform  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    pipe, geom = bench.build_pipeline(dev)
    vols_h, verts_h = bench.make_inputs(0)
    vol = torch.from_numpy(vols_h[0]).to(dev)
    verts = torch.from_numpy(verts_h).to(dev)
    acc = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        acc.setdefault(name, []).append(e0.elapsed_time(e1))
        return r

    for it in range(iters + 2):
        if it == 2:
            acc.clear()
        prob = timed("segment", lambda: pipe.segmenter.segment_device(vol, if_output_prob_map=True))
        A_r = timed("resize", lambda: (ops.resize_trilinear(vol, (80, 192, 192)),
                                        ops.resize_trilinear(pipe.atlas, (80, 192, 192))))
        phi = timed("gradicon", lambda: pipe.reg_model(*A_r))
        tr = timed("disp_field", lambda: (CompositeTransform(ops.displacement_field(phi[0][0]), geom, pipe.atlas_geom),
                                          CompositeTransform(ops.displacement_field(phi[1][0]), pipe.atlas_geom, geom)))
        timed("warp_volume", lambda: tr[0].resample_device(prob, geom, pipe.atlas_geom))
        timed("warp_points", lambda: ops.warp_points(verts, tr[1].disp, tr[1].from_network_space_inv,
                                                     tr[1].to_network_space))
        timed("whole", lambda: pipe.run_device(vol, geom, verts))
    print(json.dumps({k: round(sum(v) / len(v), 3) for k, v in acc.items()}))


if __name__ == "__main__":
    main()
