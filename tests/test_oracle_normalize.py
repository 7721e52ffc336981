"""CPU tier: the image_normalize oracle (dask_processing.py:10-26) against known answers."""
import numpy as np

from oracle.normalize_oracle import image_normalize


def test_known_answers():
    a = np.arange(1001, dtype=np.float32)                # percentiles of 0..1000 are exact: p% -> 10 p
    out, (wmin, wmax) = image_normalize(a, 10.0, 90.0, 0.0, 1.0)
    assert (wmin, wmax) == (100.0, 900.0)
    assert out[0] == 0.0 and out[100] == 0.0 and out[900] == 1.0 and out[1000] == 1.0
    assert abs(out[500] - 0.5) < 1e-7 and abs(out[300] - 0.25) < 1e-7
    assert out.dtype == np.float32


def test_linear_interpolation_between_order_statistics():
    a = np.array([0.0, 10.0, 20.0, 40.0], dtype=np.float32)   # virtual index of 50 % = 1.5 -> 15; of 90 % = 2.7 -> 34
    _, (wmin, wmax) = image_normalize(a, 50.0, 90.0, 0.0, 1.0)
    assert abs(wmin - 15.0) < 1e-6 and abs(wmax - 34.0) < 1e-5


def test_output_range_and_monotonicity():
    rng = np.random.default_rng(3)
    a = rng.normal(100.0, 30.0, 20000).astype(np.float32)
    out, (wmin, wmax) = image_normalize(a, 0.1, 99.9, -1.0, 3.0)
    assert out.min() == -1.0 and out.max() == 3.0
    o = np.argsort(a)
    assert np.all(np.diff(out[o]) >= 0)
    assert wmin < wmax


def test_percentiles_match_the_committed_numpy_fixture():
    """tests/golden/normalize_percentiles.json was written by numpy 2.3.5 (float32 arrays: virtual index and lerp in
    float32, NEP 50).  The device kernel restates exactly that; a numpy with different semantics must show up here."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "normalize_percentiles.json")
    fx = json.load(open(path))
    rng = np.random.default_rng(2024)
    for c in fx["cases"]:
        a = (rng.standard_normal(c["n"]) * 37.0 + 5.0).astype(np.float32)
        assert float(a[0]) == c["first"] and float(a[-1]) == c["last"]
        _, (wmin, wmax) = image_normalize(a, c["lo"], c["hi"], 0.0, 1.0)
        assert wmin == c["wmin"] and wmax == c["wmax"], (c, wmin, wmax)
