"""One eager knee inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oai_analysis_2_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
pipe, geom = bench.build_pipeline(dev)
vols, verts = bench.make_inputs(0)
vol, v = torch.from_numpy(vols[0]).to(dev), torch.from_numpy(verts).to(dev)
raw = vol * 900.0 + 17.0
for _ in range(2):
    ops.intensity_window(raw, 0.1, 99.9, 0.0, 1.0)
    pipe.run_device(vol, geom, v)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.intensity_window(raw, 0.1, 99.9, 0.0, 1.0)
pipe.run_device(vol, geom, v)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
