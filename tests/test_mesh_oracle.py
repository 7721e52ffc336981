"""CPU tier: the marching-cubes oracle (generated polygon table, asymptotic-decider faces) against closed forms, and
the library's independently generated table (oai_mc_table, C++) against the oracle's, entry by entry."""
import numpy as np

from oai_analysis_2_b200 import ops
from oracle import mesh_oracle as mo


def _sphere(n=40, r=12.0, c=(19.3, 20.1, 19.7)):
    z, y, x = np.meshgrid(*(np.arange(n),) * 3, indexing="ij")
    d = np.sqrt((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2)
    return 1.0 / (1.0 + np.exp(d - r))     # a probability map: high inside


def test_library_table_equals_oracle_table():
    t = ops.mc_table()
    for mask in range(256):
        for bits in range(64):
            ref = mo.table_entry(mask, bits)
            n = int(t[mask, bits, 0])
            got = [tuple(int(v) for v in t[mask, bits, 1 + 3 * i:4 + 3 * i]) for i in range(n)]
            assert got == [tuple(r) for r in ref], (mask, bits)
    assert t[0, :, 0].max() == 0 and t[255, :, 0].max() == 0 and t[:, :, 0].max() <= 10


def test_every_table_entry_is_a_set_of_closed_oriented_loops():
    for mask in range(1, 255):
        amb = [f for f in range(6) if mo.face_is_ambiguous(mask, f)]
        for sub in range(1 << len(amb)):
            bits = sum(1 << amb[i] for i in range(len(amb)) if (sub >> i) & 1)
            loops = mo.cell_polygons(mask, bits)
            crossed = {e for e, (a, b) in enumerate(mo.EDGE_CORNERS) if ((mask >> a) & 1) != ((mask >> b) & 1)}
            assert sorted(e for lp in loops for e in lp) == sorted(crossed)     # every crossed edge exactly once
            assert all(len(lp) >= 3 for lp in loops)
    # complementary patterns with complementary face decisions give the same loops, reversed
    assert mo.cell_polygons(0b00000001, 0) == [[0, 4, 8]] or len(mo.cell_polygons(0b00000001, 0)[0]) == 3


def test_sphere_area_volume_and_topology():
    vol = _sphere()
    v, f = mo.marching_cubes(vol, 0.5, (1, 1, 1), "descent")
    st = mo.mesh_stats(v, f)
    assert st["euler"] == 2
    assert abs(st["area"] - 4 * np.pi * 144) / (4 * np.pi * 144) < 0.01
    assert abs(st["volume"] - 4 / 3 * np.pi * 12 ** 3) / (4 / 3 * np.pi * 12 ** 3) < 0.01
    va, fa = mo.marching_cubes(vol, 0.5, (1, 1, 1), "ascent")           # ascent = the reference's setting: flipped winding
    assert np.allclose(va, v) and mo.mesh_stats(va, fa)["volume"] < 0
    # spacing scales the coordinates per axis
    vs, _ = mo.marching_cubes(vol, 0.5, (0.36, 0.36, 0.7), "ascent")
    assert np.allclose(vs, v * np.array([0.36, 0.36, 0.7]))
    # every vertex sits on a lattice edge, at the linear crossing of the level
    frac = v - np.floor(v)
    assert ((frac > 1e-12).sum(axis=1) <= 1).all()


def test_noise_surface_is_closed_and_consistently_oriented():
    rng = np.random.default_rng(0)
    vol = rng.random((12, 13, 14))
    vol[0], vol[-1], vol[:, 0], vol[:, -1], vol[:, :, 0], vol[:, :, -1] = 0, 0, 0, 0, 0, 0   # closed inside the volume
    v, f = mo.marching_cubes(vol, 0.5)
    d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = d[:, 0] * (len(v) + 1) + d[:, 1]
    rev = d[:, 1] * (len(v) + 1) + d[:, 0]
    # watertight and oriented: every directed edge is matched by its reverse the same number of times
    ku, kc = np.unique(key, return_counts=True)
    ru, rc = np.unique(rev, return_counts=True)
    assert np.array_equal(ku, ru) and np.array_equal(kc, rc)


def test_small_regions_are_dropped():
    vol = np.maximum(_sphere(40, 12.0), _sphere(40, 2.5, (5.2, 5.1, 5.3)))     # a big and a tiny blob
    v, f = mo.marching_cubes(vol, 0.5)
    kv, kf, nreg = mo.keep_large_regions(v, f, 3000)
    assert nreg == 1 and 3000 < len(kf) < len(f) and kf.max() == len(kv) - 1
    assert mo.mesh_stats(kv, kf)["euler"] == 2
    assert len(mo.keep_large_regions(v, f, 10 ** 6)[1]) == 0
