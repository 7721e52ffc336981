"""CPU tier: the host-side affine algebra of oai_analysis_2_b200.transforms (what feeds oai_warp_volume /
oai_warp_points) against the float64 oracle of the ITK semantics (oracle/warp_oracle.py), on oblique geometries."""
import numpy as np
import torch

from oai_analysis_2_b200 import transforms as T
from oracle import warp_oracle as O


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    return q * np.sign(np.linalg.det(q))


def _apply(a, p):
    return p @ a[0].T + a[1]


def test_geometry_and_resampling_affines_match_oracle():
    rng = np.random.default_rng(0)
    for _ in range(5):
        size = rng.integers(5, 40, 3)
        sp, org, dr = rng.uniform(0.3, 1.5, 3), rng.uniform(-50, 50, 3), _rot(rng)
        g, go = T.Geometry(size, sp, org, dr), O.Geometry(size, sp, org, dr)
        idx = rng.uniform(-2, 40, (64, 3))
        phys = go.index_to_physical(idx)
        assert np.allclose(_apply(g.index_to_physical_affine(), idx), phys, atol=1e-10)
        assert np.allclose(_apply(g.physical_to_index_affine(), phys), idx, atol=1e-9)
        net = np.array([19, 24, 11])
        M, cf, cm = O.resampling_transform(go, net)
        x = rng.uniform(0, 20, (64, 3))
        assert np.allclose(_apply(T.resampling_transform(g, net), x), (x - cf) @ M.T + cm, atol=1e-9)


def test_zero_field_composite_is_the_affine_chain_of_the_oracle():
    """With a zero displacement field T(p) = R_A(R_B^-1(p)); the two affines handed to the kernels must compose to the
    oracle's TransformPoint, and the index-space chains used by resample_device to the oracle's resample mapping."""
    rng = np.random.default_rng(1)
    gA = (rng.integers(20, 40, 3), rng.uniform(0.3, 1.0, 3), rng.uniform(-20, 20, 3), _rot(rng))
    gB = (rng.integers(20, 40, 3), rng.uniform(0.3, 1.0, 3), rng.uniform(-20, 20, 3), _rot(rng))
    disp = np.zeros((9, 12, 10, 3))
    tr = T.CompositeTransform(torch.zeros(9, 12, 10, 3), T.Geometry(*gA), T.Geometry(*gB))
    oracle = O.CompositeTransform(disp, O.Geometry(*gA), O.Geometry(*gB))
    p = O.Geometry(*gB).index_to_physical(rng.uniform(0, 20, (128, 3)))
    got = _apply(tr.to_network_space, _apply(tr.from_network_space_inv, p))
    assert np.allclose(got, oracle.transform_points(p), atol=1e-9)
    # resample_device's affines: output index (grid B) -> lattice, lattice -> source index (grid A)
    a = T._compose(tr.from_network_space_inv, T.Geometry(*gB).index_to_physical_affine())
    b = T._compose(T.Geometry(*gA).physical_to_index_affine(), tr.to_network_space)
    j = rng.integers(0, 20, (128, 3)).astype(np.float64)
    want = O.Geometry(*gA).physical_to_index(oracle.transform_points(O.Geometry(*gB).index_to_physical(j)))
    assert np.allclose(_apply(b, _apply(a, j)), want, atol=1e-8)


def test_invert_and_compose_are_consistent():
    rng = np.random.default_rng(2)
    a = (rng.standard_normal((3, 3)) + 3 * np.eye(3), rng.standard_normal(3))
    b = (rng.standard_normal((3, 3)) + 3 * np.eye(3), rng.standard_normal(3))
    x = rng.standard_normal((16, 3))
    assert np.allclose(_apply(T._compose(a, b), x), _apply(a, _apply(b, x)))
    assert np.allclose(_apply(T._invert(a), _apply(a, x)), x)
