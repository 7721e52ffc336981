"""GPU tier: oai_intensity_window (device radix select + windowing) against the numpy oracle of
dask_processing.image_normalize.  The order statistics are exact, so the window must match np.percentile to float32
rounding and the output to one float32 ulp of the rescale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _check(a, lo, hi, omin, omax, tol=2e-6):
    from oai_analysis_2_b200 import ops
    from oracle.normalize_oracle import image_normalize
    ref, (wmin, wmax) = image_normalize(a, lo, hi, omin, omax)
    got, (gmin, gmax) = ops.intensity_window(torch.from_numpy(a).cuda(), lo, hi, omin, omax, return_window=True)
    got = got.cpu().numpy()
    scale = max(abs(wmin), abs(wmax), 1e-30)
    assert abs(gmin - wmin) <= 2e-7 * scale and abs(gmax - wmax) <= 2e-7 * scale, (gmin, wmin, gmax, wmax)
    span = abs(omax - omin)
    # voxels within an ulp of the window edges may fall on either side when the window itself differs by an ulp
    bad = np.abs(got - ref) > tol * span
    assert bad.sum() == 0, (bad.sum(), np.abs(got - ref).max())
    return gmin, gmax


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4099, 1 << 20])
def test_window_matches_numpy_random(n):
    _cuda()
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) * 37.0 + 5.0).astype(np.float32)
    if n >= 2:
        _check(a, 0.1, 99.9, 0.0, 1.0)
        _check(a, 25.0, 75.0, -2.0, 5.0)
    else:
        from oai_analysis_2_b200 import ops
        _, (gmin, gmax) = ops.intensity_window(torch.from_numpy(a).cuda(), 0.1, 99.9, 0.0, 1.0, return_window=True)
        assert gmin == gmax == float(a[0])       # a flat image has an empty window, as in the reference


def test_window_with_negative_values_ties_and_extremes():
    _cuda()
    rng = np.random.default_rng(0)
    a = np.concatenate([np.full(5000, -3.5, np.float32), np.zeros(3000, np.float32), -np.zeros(10, np.float32),
                        rng.integers(-50, 50, 20000).astype(np.float32), np.array([1e30, -1e30, 1e-30], np.float32)])
    rng.shuffle(a)
    _check(a, 0.0, 100.0, 0.0, 1.0, tol=1e-5)
    _check(a, 1.0, 99.0, 0.0, 255.0)


def test_full_size_knee_and_host_mirror():
    _cuda()
    from oai_analysis_2_b200 import dask_processing, synthetic
    from oracle.normalize_oracle import image_normalize
    vol = (synthetic.synthetic_knee((160, 384, 384), seed=5) * 900.0 + 17.0).astype(np.float32)   # raw-DESS-like range
    gmin, gmax = _check(vol.reshape(-1), 0.1, 99.9, 0.0, 1.0)
    assert 17.0 <= gmin < gmax <= 917.0
    out = dask_processing.image_normalize(vol, 0.1, 99.9, 0, 1)
    ref, _ = image_normalize(vol, 0.1, 99.9, 0, 1)
    assert np.asarray(out).shape == vol.shape and np.abs(np.asarray(out) - ref).max() <= 2e-6
    # idempotence on the already-windowed volume: the percentiles of the output are the output range ends
    again = dask_processing.image_normalize(np.asarray(out), 0.1, 99.9, 0, 1)
    assert np.abs(np.asarray(again) - np.asarray(out)).max() <= 2e-6
