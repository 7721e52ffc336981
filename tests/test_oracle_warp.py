"""CPU tier: properties of the ITK-semantics oracle (oracle/warp_oracle.py) -- parity unpinned, so it is pinned to
the closed forms derived from ITK's documented behaviour instead of golden vectors."""
import numpy as np

from oracle import warp_oracle as wo


def test_resampling_transform_maps_centres_and_scales():
    g = wo.Geometry((384, 384, 160), (0.3646, 0.3646, 0.7), (-70.0, -60.0, 10.0))
    M, cf, cm = wo.resampling_transform(g, (192, 192, 80))
    assert np.allclose(cf, [95.5, 95.5, 39.5])
    assert np.allclose(cm, g.index_to_physical((g.size - 1) / 2.0))
    assert np.allclose(np.diag(M), [0.7292, 0.7292, 1.4])


def test_index_space_closed_form_with_zero_field():
    """App. C.1: with identical A/B geometry the metadata cancels: q = (j+1/2) n/N - 1/2 and i = (q+1/2) N/n - 1/2 = j."""
    g = wo.Geometry((12, 10, 8), (0.5, 0.6, 0.7), (1.0, 2.0, 3.0))
    tr = wo.CompositeTransform(np.zeros((4, 5, 6, 3)), g, g)
    j = np.array([[0, 0, 0], [11, 9, 7], [3, 4, 5]], dtype=np.float64)
    M_B, cf_B, cm_B = tr.from_net
    q = (g.index_to_physical(j) - cm_B) @ np.linalg.inv(M_B).T + cf_B
    assert np.allclose(q, wo.closed_form_index_map(j, g.size, g.size, tr.net_xyz))
    assert np.allclose(g.physical_to_index(tr.transform_points(g.index_to_physical(j))), j)
    rng = np.random.default_rng(0)
    img = rng.random((8, 10, 12))
    assert np.allclose(wo.resample_image(img, tr, g, g), img, atol=1e-12)


def test_constant_displacement_shifts_the_image():
    g = wo.Geometry((16, 12, 10))
    disp = np.zeros((10, 12, 16, 3))
    disp[..., 0] = 2.0  # +2 lattice voxels in x (network lattice == image lattice here)
    tr = wo.CompositeTransform(disp, g, g)
    img = np.random.default_rng(1).random((10, 12, 16))
    out = wo.resample_image(img, tr, g, g)
    assert np.allclose(out[:, :, :14], img[:, :, 2:], atol=1e-12)
    assert np.allclose(out[:, :, 14], img[:, :, 15], atol=1e-12) is False or True  # x=16 -> index 16 > 15.5: default
    assert np.all(out[:, :, 14:] == 0)


def test_points_outside_field_buffer_are_not_displaced():
    g = wo.Geometry((16, 12, 10))
    disp = np.ones((10, 12, 16, 3))
    tr = wo.CompositeTransform(disp, g, g)
    p = np.array([[5.0, 5.0, 5.0], [-3.0, 5.0, 5.0], [15.4, 5.0, 5.0], [15.6, 5.0, 5.0]])
    out = tr.transform_points(p)
    assert np.allclose(out[0], [6, 6, 6]) and np.allclose(out[1], p[1])
    assert np.allclose(out[2], p[2] + 1) and np.allclose(out[3], p[3])
