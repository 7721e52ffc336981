"""Deterministic synthetic inputs for benchmarks and smoke runs: there is no network for the OAI
release data, so knees are smooth blob fields with noise in [0,1], and thickness-mesh vertices are uniform samples of
the image's physical bounding box shrunk by 10 %."""
import numpy as np

OAI_SPACING = (0.3646, 0.3646, 0.7)  # x, y, z mm: typical OAI DESS voxel size
OAI_SHAPE = (160, 384, 384)          # z, y, x (notebooks/FullDemo.ipynb:280-281)
N_VERTS_FC, N_VERTS_TC = 64900, 20470  # reference test/test_all.py:69-70


def synthetic_knee(shape_zyx=OAI_SHAPE, seed=0, n_blobs=64):
    rng = np.random.default_rng(seed)
    D, H, W = shape_zyx
    z = np.arange(D, dtype=np.float32)[:, None, None]
    y = np.arange(H, dtype=np.float32)[None, :, None]
    x = np.arange(W, dtype=np.float32)[None, None, :]
    vol = np.zeros(shape_zyx, dtype=np.float32)
    scale = min(shape_zyx) / 160.0
    for _ in range(n_blobs):
        c = rng.uniform(0, 1, 3) * np.array(shape_zyx)
        s = rng.uniform(6, 30) * max(scale, 0.15)
        a = rng.uniform(0.3, 1.0)
        vol += a * (np.exp(-((z - c[0]) ** 2) / (2 * s * s)) * np.exp(-((y - c[1]) ** 2) / (2 * s * s))
                    * np.exp(-((x - c[2]) ** 2) / (2 * s * s))).astype(np.float32)
    vol += 0.05 * rng.uniform(0, 1, shape_zyx).astype(np.float32)
    vol -= vol.min()
    vol /= vol.max()
    return vol.astype(np.float32)


def synthetic_vertices(n, shape_zyx=OAI_SHAPE, spacing_xyz=OAI_SPACING, origin_xyz=(0, 0, 0), seed=0):
    rng = np.random.default_rng(seed)
    size = np.array(shape_zyx[::-1], dtype=np.float64)
    ext = (size - 1) * np.asarray(spacing_xyz)
    lo, hi = 0.05 * ext, 0.95 * ext
    return np.asarray(origin_xyz, dtype=np.float64) + rng.uniform(lo, hi, (n, 3))
