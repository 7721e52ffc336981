"""CPU tier: the synthetic-batch generator of scripts/run_batch.py is a pure function of the knee index."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def test_knee_volume_is_deterministic_and_distinct():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "run_batch.py")
    src = open(path).read()
    ns = {}
    start = src.index("def knee_volume")
    exec("import numpy as np\n" + src[start:src.index("def main")], ns)   # the function only, no CUDA imports
    rng = np.random.default_rng(0)
    bases = [rng.random((4, 6, 16)).astype(np.float32) for _ in range(2)]
    vols = [ns["knee_volume"](bases, i) for i in range(12)]
    again = [ns["knee_volume"](bases, i) for i in range(12)]
    assert all(np.array_equal(a, b) for a, b in zip(vols, again))
    assert all(v.flags["C_CONTIGUOUS"] and v.shape == (4, 6, 16) for v in vols)
    assert len({v.tobytes() for v in vols}) == 12
