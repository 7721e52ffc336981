"""CPU tier: the 8f-4 oracles against the libraries the reference itself calls (sklearn KernelPCA, scipy leastsq) and
against closed forms."""
import numpy as np

from oracle import mesh_oracle as mo


def test_linear_pca_oracle_equals_sklearn_kernel_pca():
    from sklearn.decomposition import KernelPCA
    rng = np.random.default_rng(0)
    for n, scale in ((150, (5, 2, 0.5)), (400, (1, 9, 3))):      # dense and arpack eigensolver paths of sklearn
        x = rng.normal(size=(n, 3)) * scale + (40, 30, 70)
        want = KernelPCA(n_components=2, degree=3.0).fit_transform(x)
        assert np.abs(mo.linear_kernel_pca2(x) - want).max() < 1e-9


def test_circle_fit_recovers_an_exact_arc():
    th = np.linspace(0.1, 2.2, 300)
    x, y = 3.0 + 40.0 * np.cos(th), -2.0 + 40.0 * np.sin(th)
    c, r = mo.compute_least_square_circle(x, y)
    assert np.allclose(c, (3.0, -2.0), atol=1e-8) and abs(r - 40.0) < 1e-8


def test_map_attributes_averages_in_radius_and_falls_back_to_the_closest_point():
    src = np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.9, 0], [10, 0, 0]], dtype=np.float64)
    attr = np.array([1.0, 2.0, 6.0, 100.0])
    tgt = np.array([[0, 0, 0], [0.5, 0, 0], [5.3, 0, 0], [1.0, 0, 0]])
    got = mo.map_attributes(src, attr, tgt, radius=1.0)
    # (0,0,0): all three near points; (0.5,0,0): the third is sqrt(0.25+0.81) > 1 away; (5.3,0,0): nothing in range,
    # closest is x = 10; (1,0,0): distance exactly 1 to the origin counts (<=), 0.5 to the second
    assert np.allclose(got, [3.0, 1.5, 100.0, 1.5])


def test_tibial_projection_layout():
    rng = np.random.default_rng(1)
    left = rng.normal(size=(200, 3)) * (10, 6, 3) + (40, 30, 25)
    right = rng.normal(size=(150, 3)) * (10, 6, 3) + (40, 30, 75)
    v = np.concatenate((left, right))
    th = np.arange(350.0)
    x, y, t = mo.project_thickness_tc(v, th)
    assert len(x) == 350 and np.array_equal(t[:150], th[200:]) and np.array_equal(t[150:], th[:200])
    assert abs(y[:150].mean() - 50) < 1e-9 and abs(y[150:].mean()) < 1e-9 and abs(x.mean()) < 1e-9
