"""GPU tier: atlas attribute mapping and 2-D thickness projection against the oracles.  The circle fit
is checked against the reference's own scipy.optimize.leastsq call and the tibial flattening against sklearn's
KernelPCA (both libraries are installed; vtk is not, so map_attributes' oracle is a restatement)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _femoral_sheet(n=6000, seed=0):
    """points on ~140 degrees of a noisy cylinder along z (what the femoral cartilage looks like to get_cylinder)"""
    rng = np.random.default_rng(seed)
    th = rng.uniform(-0.2, 2.3, n)
    r = 38.0 + rng.normal(0, 0.6, n)
    z = rng.uniform(20.0, 95.0, n)
    # the reference swaps x and y before fitting: build the cylinder in (y, x)
    return np.stack((12.0 + r * np.sin(th), -7.0 + r * np.cos(th), z), axis=1).astype(np.float32)


def _tibial_plateaus(n=5000, seed=1):
    rng = np.random.default_rng(seed)
    out = []
    for zc, tilt in ((28.0, 0.3), (74.0, -0.4)):
        u, v = rng.uniform(-15, 15, n), rng.uniform(-9, 9, n)
        w = 0.02 * u * u + rng.normal(0, 0.2, n)
        out.append(np.stack((40 + u * np.cos(tilt) - w * np.sin(tilt), 30 + u * np.sin(tilt) + w * np.cos(tilt), zc + v),
                            axis=1))
    return np.concatenate(out).astype(np.float32)


@pytest.mark.parametrize("n_attr", [1, 3])
def test_map_attributes_matches_oracle(n_attr):
    _cuda()
    from oai_analysis_2_b200 import mesh_processing as mp
    from oracle import mesh_oracle as mo
    rng = np.random.default_rng(3)
    src = rng.uniform(0, 30, (4000, 3)).astype(np.float32)
    tgt = np.concatenate((rng.uniform(0, 30, (1500, 3)), rng.uniform(40, 45, (40, 3)))).astype(np.float32)  # + far points
    attr = rng.uniform(0.5, 4.0, (4000,) if n_attr == 1 else (4000, n_attr)).astype(np.float32)
    want = mo.map_attributes(src, attr, tgt, radius=1.0)
    sv, sa, tv = (torch.from_numpy(a).cuda() for a in (src, attr, tgt))
    faces = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    v, f, got = mp.map_attributes((sv, faces, sa), (tv, faces))
    assert v is tv and got.shape == tuple(want.shape)
    # the in-range test is d2 <= r2 in float32 on the device and float64 in the oracle: exclude the (measure-zero)
    # targets with a source point within 1e-5 of the radius
    d = np.sqrt(((tgt[:, None, :].astype(np.float64) - src[None]) ** 2).sum(-1))
    safe = ~(np.abs(d - 1.0) < 1e-5).any(1)
    err = np.abs(got.cpu().numpy().astype(np.float64) - want)[safe]
    print(f"map_attributes[{n_attr}]: {safe.sum()} of {len(tgt)} targets compared, max |err| {err.max():.2e}; "
          f"null points {(d.min(1) > 1.0).sum()}")
    assert err.max() < 2e-6 * 4.0
    assert (d.min(1) > 1.0).sum() >= 40      # the far targets exercised the closest-point rule


def test_femoral_projection_matches_the_reference_scipy_fit():
    _cuda()
    from oai_analysis_2_b200 import mesh_processing as mp, ops
    from oracle import mesh_oracle as mo
    pts = _femoral_sheet()
    th = np.random.default_rng(5).uniform(1, 3, len(pts)).astype(np.float32)
    angle_w, z_w, th_w, center_w = mo.project_thickness_fc(pts, th)
    v = torch.from_numpy(pts).cuda()
    center, radius, iters = ops.circle_fit(v, 1, 0)
    print(f"circle fit: centre {center} vs scipy {tuple(center_w)}; R {radius:.4f}; {iters} Gauss-Newton steps")
    assert np.abs(np.array(center) - center_w).max() < 1e-6
    x, y, t = mp.project_thickness((v, None, torch.from_numpy(th).cuda()), "FC")
    assert x.dtype == torch.float64
    assert np.abs(x.cpu().numpy() - angle_w).max() < 1e-7
    assert np.array_equal(y.cpu().numpy(), z_w) and np.array_equal(t.cpu().numpy(), th_w)
    (c2, r2), (z0, z1) = mp.get_cylinder(v)
    assert abs(r2 - 38.0) < 0.2 and z0 == pts[:, 2].min() and z1 == pts[:, 2].max()


def test_tibial_projection_matches_sklearn_kernel_pca():
    _cuda()
    from sklearn.decomposition import KernelPCA
    from oai_analysis_2_b200 import mesh_processing as mp, ops
    from oracle import mesh_oracle as mo
    pts = _tibial_plateaus()
    th = np.random.default_rng(6).uniform(1, 3, len(pts)).astype(np.float32)
    v = torch.from_numpy(pts).cuda()
    # the flattening itself, against sklearn's own KernelPCA on a subset small enough for its n x n Gram matrix
    sub = np.arange(0, 5000, 5)
    want = KernelPCA(n_components=2, degree=3.0).fit_transform(pts[sub])
    gx, gy = ops.pca2_project(v, torch.from_numpy(sub).cuda().int())
    got = np.stack((gx.cpu().numpy(), gy.cpu().numpy()), 1)
    print(f"KernelPCA scores: max |err| {np.abs(got - want).max():.2e} on scores up to {np.abs(want).max():.1f}")
    assert np.abs(got - want).max() < 5e-5      # sklearn itself works in float32 on float32 input (arpack start vector is random)
    # the whole tibial branch against the oracle (same algebra as the reference, PCA through the covariance)
    xw, yw, tw = mo.project_thickness_tc(pts, th)
    x, y, t = mp.project_thickness((v, None, torch.from_numpy(th).cuda()), "TC")
    assert np.abs(x.cpu().numpy() - xw).max() < 1e-5 and np.abs(y.cpu().numpy() - yw).max() < 1e-5
    assert np.array_equal(t.cpu().numpy(), tw)
    assert (y.cpu().numpy()[:5000].mean() > 40) and abs(y.cpu().numpy()[5000:].mean()) < 1   # right plateau lifted by 50


def test_thickness_to_atlas_to_projection_to_vtk_file(tmp_path):
    """The tail of the reference's per-knee chain in one go (mesh_processing.py:381-534 + itk.meshwrite): thickness mesh
    of a slab -> mapped onto an 'atlas' surface (the same sheet, shifted and decimated) -> 2-D projection -> .vtk file."""
    _cuda()
    from test_mesh_gpu import _slab_volume
    from oai_analysis_2_b200 import io as oio, itk_compat, mesh_processing as mp
    sp = (0.36, 0.36, 0.7)
    th = mp.get_thickness_mesh(itk_compat.Image(_slab_volume(), spacing=sp), mesh_type="TC")
    iv, if_, d_in = th["inner"]
    assert d_in.shape[0] == iv.shape[0] and float(d_in.median()) > 1.0   # 0 at the rim where the two surfaces meet
    # atlas = every third vertex of the inner surface, nudged by a fraction of the mapping radius
    g = torch.Generator(device="cuda").manual_seed(0)
    atlas_v = (iv[::3] + 0.05 * torch.randn(iv[::3].shape, device="cuda", generator=g)).contiguous()
    atlas_f = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    mv, mf, mapped = mp.map_attributes((iv, if_, d_in), (atlas_v, atlas_f))
    assert mapped.shape == (atlas_v.shape[0],)
    # a radius-1 mm average of a smooth thickness field stays close to the thickness of the vertex it sits on
    dev = (mapped - d_in[::3]).abs()
    print(f"mapped thickness vs the vertex's own: median |diff| {float(dev.median()):.3f} mm, max {float(dev.max()):.3f} mm "
          f"(thickness {float(d_in.median()):.2f} mm)")
    assert float(dev.median()) < 0.1 * float(d_in.median())
    # femoral-style unrolling of the mapped mesh (any sheet can be unrolled around its best-fit cylinder)
    ang, hgt, val = mp.project_thickness((mv, mf, mapped), "FC")
    assert ang.shape == hgt.shape == val.shape and float(ang.abs().max()) <= np.pi + 1e-9
    # tibial-style flattening needs the two plateaus either side of z = 50 mm: lift a copy of the sheet
    two = torch.cat((mv, mv + torch.tensor([0.0, 0.0, 60.0], device="cuda")))
    x2, y2, t2 = mp.project_thickness((two, mf, torch.cat((mapped, mapped))), "TC")
    n = mv.shape[0]
    assert x2.shape[0] == 2 * n and abs(float(y2[:n].mean()) - 50.0) < 1e-6 and abs(float(y2[n:].mean())) < 1e-6
    # the two plateaus are congruent: same PCA scores up to the two rotations and the mirror
    assert abs(float(x2[:n].std()) ** 2 + float(y2[:n].std()) ** 2 - float(x2[n:].std()) ** 2 - float(y2[n:].std()) ** 2) < 1e-4
    path = str(tmp_path / "inner_thickness.vtk")
    oio.write_vtk_mesh(path, iv, if_, {"thickness": d_in}, binary=True)
    rv, rf, rd = oio.read_vtk_mesh(path)
    assert np.array_equal(rv, iv.cpu().numpy()) and np.array_equal(rf, if_.cpu().numpy())
    assert np.array_equal(rd["thickness"], d_in.cpu().numpy())
