#!/usr/bin/env python
"""BASELINE configs[3]: a batch of synthetic knees sharded by volume over the GPUs of one box.

    python scripts/run_batch.py --knees 64                                   # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_batch.py --knees 64

Knee i goes to rank i % N (static round-robin, `sharding.shard_indices`); every rank streams its knees through the
captured per-knee CUDA graph (`KneePipeline.run_stream`: H2D / compute / D2H overlapped); the only communication is a
host-side gather of one small record per knee (checksums of the atlas-space maps and warped vertices) -- no data-path
collective.  Rank 0 prints one JSON line; the per-knee checksums do not depend on N (compare runs with --dump).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oai_analysis_2_b200 import sharding, synthetic  # noqa: E402


def knee_volume(bases, i):
    """Deterministic, cheap-to-make knee number i: a mirrored / rolled variant of one of a few blob fields."""
    b = bases[i % len(bases)]
    v = b[:, ::-1, :] if (i // len(bases)) % 2 else b
    return np.ascontiguousarray(np.roll(v, 7 * (i // (2 * len(bases))) + 3 * i, axis=2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--knees", type=int, default=64)
    ap.add_argument("--dump", default=None, help="write the gathered per-knee records to this JSON file (rank 0)")
    args = ap.parse_args()
    rank, world, local = sharding.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pipe, geom = bench.build_pipeline(dev)
    verts = np.concatenate([synthetic.synthetic_vertices(synthetic.N_VERTS_FC, seed=0),
                            synthetic.synthetic_vertices(synthetic.N_VERTS_TC, seed=1000)])
    pipe.capture(synthetic.OAI_SHAPE, geom, verts.shape[0])
    bases = [synthetic.synthetic_knee(synthetic.OAI_SHAPE, seed=200 + k) for k in range(2)]
    mine = sharding.shard_indices(args.knees, rank, world)
    pins = [torch.from_numpy(knee_volume(bases, i)).pin_memory() for i in mine[:4]]   # 4 pinned slots, refilled
    vp = torch.from_numpy(verts).pin_memory()

    def items():
        for k, i in enumerate(mine):
            slot = pins[k % len(pins)]
            if k >= len(pins):
                slot.copy_(torch.from_numpy(knee_volume(bases, i)))   # host-side producer (would be the image reader)
            yield slot, vp

    for _ in pipe.run_stream((pins[k % len(pins)], vp) for k in range(2)):   # warm-up, allocates the stream buffers
        pass
    sharding.barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    recs = []
    for i, res in zip(mine, pipe.run_stream(items())):
        recs.append(dict(index=i, rank=rank,
                         fc_sum=float(res["FC_atlas"].sum(dtype=np.float64)),
                         tc_sum=float(res["TC_atlas"].sum(dtype=np.float64)),
                         vert_sum=float(res["vertices_atlas"].sum()),
                         field_abs_max=float(np.abs(res["phi_AB_field"]).max())))
    torch.cuda.synchronize()
    dt = sharding.max_over_ranks(time.perf_counter() - t0, world)
    merged = sharding.gather_records(recs, rank, world)
    if rank == 0:
        assert [r["index"] for r in merged] == list(range(args.knees))
        if args.dump:
            with open(args.dump, "w") as f:
                json.dump(merged, f)
        print(json.dumps(dict(config="BASELINE configs[3]: batch of knees sharded by volume", knees=args.knees,
                              n_gpus=world, seconds=dt, knees_per_s=args.knees / dt,
                              checksum=float(sum(r["fc_sum"] + r["tc_sum"] + r["vert_sum"] for r in merged)))),
              flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
