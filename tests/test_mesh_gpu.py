"""GPU tier: mesh post-processing against the oracles -- smoothing, face features, closest-point
distance, the 2-means split against the reference's own sklearn call, and the thickness chain on a slab of known
thickness."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _slab_volume(shape=(40, 72, 64), thickness=6.0, sharp=2.0):
    """a curved cartilage-like sheet of constant thickness (voxels), probability high inside"""
    z, y, x = np.meshgrid(*(np.arange(n, dtype=np.float32) for n in shape), indexing="ij")
    mid = 0.45 * shape[1] + 6.0 * np.sin(x / shape[2] * np.pi) + 0.08 * (z - shape[0] / 2) ** 2 / 4
    d = np.abs(y - mid) - thickness / 2
    edge = np.minimum(np.minimum(x - 6, shape[2] - 7 - x), np.minimum(z - 5, shape[0] - 6 - z))
    d = np.maximum(d, -edge)
    return (1.0 / (1.0 + np.exp(sharp * d))).astype(np.float32)


def _mesh(spacing=(0.36, 0.36, 0.7)):
    from oai_analysis_2_b200 import mesh_processing as mp
    vol = torch.from_numpy(_slab_volume()).cuda()
    return mp.extract_isosurface_device(vol, spacing)


def test_smoothing_matches_oracle():
    _cuda()
    from oai_analysis_2_b200 import ops
    from oracle import mesh_oracle as mo
    verts, faces = _mesh()
    v, f = verts.cpu().numpy(), faces.cpu().numpy().astype(np.int64)
    keep = 700                         # the python oracle walks vertex by vertex: a patch is enough
    sub = f[(f < keep).all(axis=1)]
    got = ops.smooth_mesh(torch.from_numpy(v[:keep]).cuda(), torch.from_numpy(sub.astype(np.int32)).cuda(), 30, 0.01)
    jac = mo.smooth_mesh(v[:keep], sub, 30, 0.01, in_place=False)
    gs = mo.smooth_mesh(v[:keep], sub, 30, 0.01, in_place=True)
    moved = np.abs(jac - v[:keep]).max()
    assert moved > 1e-3
    assert np.abs(got.cpu().numpy() - jac).max() < 2e-6                      # same arithmetic: float32 steps
    assert np.abs(got.cpu().numpy() - gs).max() < 0.05 * moved               # VTK's in-place sweep differs at O(f^2)
    # open-boundary vertices of the patch stayed put
    _, fixed = mo.vertex_neighbours(keep, sub)
    assert fixed.any() and np.array_equal(got.cpu().numpy()[fixed], v[:keep][fixed])
    # the full closed mesh: 150 sweeps shrink it slightly and keep it finite
    full = ops.smooth_mesh(verts, faces, 150, 0.01)
    assert torch.isfinite(full).all() and float((full - verts).abs().max()) < 1.0


def test_face_features_and_distance_match_oracle():
    _cuda()
    from oai_analysis_2_b200 import ops
    from oracle import mesh_oracle as mo
    verts, faces = _mesh()
    v, f = verts.cpu().numpy(), faces.cpu().numpy().astype(np.int64)
    n, c = ops.face_features(verts, faces)
    rn, rc = mo.face_normals_centroids(v, f)
    assert np.abs(n.cpu().numpy() - rn).max() < 2e-5 and np.abs(c.cpu().numpy() - rc).max() < 1e-5
    rng = np.random.default_rng(0)
    pts = (v[rng.integers(0, len(v), 300)] + rng.normal(0, 1.5, (300, 3))).astype(np.float32)
    sub = f[::7]
    d = ops.mesh_distance(torch.from_numpy(pts).cuda(), verts, torch.from_numpy(sub.astype(np.int32)).cuda())
    ref = mo.point_mesh_distance(pts, v, sub)
    assert np.abs(d.cpu().numpy() - ref).max() < 1e-4
    # a point on a triangle has distance 0; a point lifted along the normal by h has distance <= h
    tri = v[f[10]]
    p0 = tri.mean(0).astype(np.float32)
    d0 = ops.mesh_distance(torch.from_numpy(p0[None]).cuda(), verts, faces)
    assert float(d0[0]) < 1e-5


def test_kmeans_split_matches_sklearn_and_thickness_of_a_slab():
    _cuda()
    from oai_analysis_2_b200 import itk_compat, mesh_processing as mp, ops
    from oracle import mesh_oracle as mo
    sp = (0.36, 0.36, 0.7)
    verts, faces = _mesh(sp)
    verts = mp.smooth_mesh(verts, faces, 150)
    normals, cent = ops.face_features(verts, faces)
    # tibial split: one clustering on [centroid_norm, 10 * normal]
    lab = mp.split_tibial_cartilage_surface(verts, faces, normals, cent).cpu().numpy()
    ref, feats = mo.split_tibial(normals.cpu().numpy().astype(np.float64), cent.cpu().numpy().astype(np.float64))
    agree = (lab == ref).mean()
    print(f"tibial split: agreement with sklearn KMeans {agree:.5f}; inner faces {int((lab == -1).sum())} outer {int((lab == 1).sum())}")
    assert agree > 0.995
    # femoral split: three x-segments, 9-D features
    labf = mp.split_femoral_cartilage_surface(verts, faces, normals, cent).cpu().numpy()
    v = verts.cpu().numpy().astype(np.float64)
    reff = mo.split_femoral(normals.cpu().numpy().astype(np.float64), cent.cpu().numpy().astype(np.float64), v.min(0), v.max(0))
    agree_f = (labf == reff).mean()
    print(f"femoral split: agreement with sklearn KMeans {agree_f:.5f}")
    assert agree_f > 0.99
    # thickness chain on the slab: 6 voxels along y = 6 * 0.36 mm (the sheet is tilted a little: allow 15 %)
    img = itk_compat.Image(_slab_volume(), spacing=sp)
    th = mp.get_thickness_mesh(img, mesh_type="TC")
    d_in, d_out = th["inner"][2].cpu().numpy(), th["outer"][2].cpu().numpy()
    assert len(d_in) > 500 and len(d_out) > 500
    med = float(np.median(np.concatenate([d_in, d_out])))
    print(f"slab thickness: median {med:.3f} mm (nominal {6 * 0.36:.3f})")
    assert abs(med - 6 * 0.36) / (6 * 0.36) < 0.15
    # distances agree with the oracle on a sample
    iv, if_ = th["inner"][0], th["inner"][1]
    ov, of = th["outer"][0], th["outer"][1]
    idx = np.arange(0, len(d_in), max(1, len(d_in) // 200))
    ref_d = mo.point_mesh_distance(iv.cpu().numpy()[idx], ov.cpu().numpy(), of.cpu().numpy().astype(np.int64))
    assert np.abs(d_in[idx] - ref_d).max() < 1e-4
