"""Writes tests/golden/normalize_percentiles.json: np.percentile (the reference's own call in
oai_analysis/dask_processing.py:16-17) on seeded float32 arrays, as computed by the numpy of this container.
Run from the repo root: python tests/golden/make_normalize_fixture.py"""
import json
import os

import numpy as np

rng = np.random.default_rng(2024)
cases = []
for n, (lo, hi) in [(1000, (0.1, 99.9)), (4099, (25.0, 75.0)), (100003, (0.1, 99.9)), (17, (10.0, 90.0))]:
    a = (rng.standard_normal(n) * 37.0 + 5.0).astype(np.float32)
    cases.append(dict(seed_note="default_rng(2024) sequential", n=n, lo=lo, hi=hi,
                      wmin=float(np.percentile(a, lo)), wmax=float(np.percentile(a, hi)),
                      first=float(a[0]), last=float(a[-1])))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "normalize_percentiles.json")
json.dump(dict(numpy=np.__version__, cases=cases), open(out, "w"), indent=1)
print("wrote", out)
