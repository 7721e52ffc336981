"""CPU oracle for the ITK side of the path -- TEST INFRASTRUCTURE ONLY.  ** parity unpinned **

float64 numpy restatement of what the reference asks ITK 5.3 to do (itk is not installable offline and the
reference's tests assert no warp number: test/test_all.py:69-70 are commented out):
  * icon_registration.itk_wrapper.resampling_transform / create_itk_transform  (composite transform
    T(p) = R_A( D( R_B^-1(p) ) ), reference call sites oai_analysis/registration.py:25, dask_processing.py:85)
  * itk.DisplacementFieldTransform.TransformPoint (linear vector interpolation, identity outside the field buffer)
  * itk.resample_image_filter(prob, transform=phi_AB, LinearInterpolateImageFunction, output grid of image_B,
    default pixel 0)  (oai_analysis/dask_processing.py:100-109, test/test_all.py:42-52)
  * itk.Transform.TransformPoint applied to mesh vertices

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np


class Geometry:
    """Image grid metadata in ITK convention (x,y,z): size, spacing, origin, direction (3x3)."""

    def __init__(self, size_xyz, spacing=(1, 1, 1), origin=(0, 0, 0), direction=None):
        self.size = np.asarray(size_xyz, dtype=np.int64)
        self.spacing = np.asarray(spacing, dtype=np.float64)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64)

    def index_to_physical(self, idx_xyz):
        return self.origin + (np.asarray(idx_xyz, dtype=np.float64) * self.spacing) @ self.direction.T

    def physical_to_index(self, p):
        return ((np.asarray(p, dtype=np.float64) - self.origin) @ np.linalg.inv(self.direction).T) / self.spacing


def resampling_transform(geom, net_shape_xyz):
    """icon itk_wrapper.resampling_transform: centred affine taking network-lattice coordinates (unit spacing, zero
    origin, size net_shape) to the physical space of `geom`:  R(x) = M (x - c_f) + c_m."""
    n = np.asarray(net_shape_xyz, dtype=np.float64)
    c_f = (n - 1) / 2.0
    c_m = geom.index_to_physical((geom.size - 1) / 2.0)
    M = geom.direction @ np.diag(geom.spacing * geom.size / n)
    return M, c_f, c_m


def _trilinear_clamped(vol, idx_zyx):
    """Linear interpolation with neighbours clamped to the buffer (ITK LinearInterpolateImageFunction /
    VectorLinearInterpolateImageFunction).  vol: [D,H,W] or [D,H,W,C]; idx: [...,3] continuous index (z,y,x)."""
    shape = np.asarray(vol.shape[:3])
    base = np.floor(idx_zyx)
    frac = idx_zyx - base
    base = base.astype(np.int64)
    out = 0.0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                off = np.array([dz, dy, dx])
                nb = np.clip(base + off, 0, shape - 1)
                w = np.prod(np.where(off == 1, frac, 1.0 - frac), axis=-1)
                v = vol[nb[..., 0], nb[..., 1], nb[..., 2]]
                out = out + (w[..., None] * v if v.ndim > w.ndim else w * v)
    return out


def _inside(idx, size):
    """ImageFunction::IsInsideBuffer: start-0.5 <= index < end+0.5 on every axis."""
    return np.all((idx >= -0.5) & (idx < np.asarray(size) - 0.5), axis=-1)


class CompositeTransform:
    """phi_AB of create_itk_transform(phi, ident, image_A, image_B): applied to points of B's physical space."""

    def __init__(self, disp_zyx3, geom_A, geom_B):
        self.disp = np.asarray(disp_zyx3, dtype=np.float64)  # [D,H,W,3], components x,y,z in network-voxel units
        d, h, w = self.disp.shape[:3]
        self.net_xyz = np.array([w, h, d])
        self.geom_A, self.geom_B = geom_A, geom_B
        self.to_net = resampling_transform(geom_A, self.net_xyz)     # network lattice -> A physical  ("to_network_space")
        self.from_net = resampling_transform(geom_B, self.net_xyz)   # its inverse is applied first

    def transform_points(self, pts_xyz):
        p = np.asarray(pts_xyz, dtype=np.float64)
        M_B, cf_B, cm_B = self.from_net
        q = (p - cm_B) @ np.linalg.inv(M_B).T + cf_B                 # R_B^-1
        inside = _inside(q, self.net_xyz)
        d = _trilinear_clamped(self.disp, q[..., ::-1])               # index the [z,y,x] array
        q = q + np.where(inside[..., None], d, 0.0)                   # DisplacementFieldTransform (identity outside)
        M_A, cf_A, cm_A = self.to_net
        return (q - cf_A) @ M_A.T + cm_A                              # R_A


def resample_image(prob_zyx, transform, geom_in, geom_out, chunk=1 << 20, default=0.0):
    """itk.resample_image_filter(prob, transform, linear interpolator, output grid = geom_out)."""
    prob = np.asarray(prob_zyx, dtype=np.float64)
    W, H, D = (int(v) for v in geom_out.size)
    out = np.empty(D * H * W, dtype=np.float64)
    for s in range(0, out.size, chunk):
        lin = np.arange(s, min(s + chunk, out.size))
        j = np.stack([lin % W, (lin // W) % H, lin // (W * H)], -1)  # x,y,z
        p = geom_out.index_to_physical(j)
        q = transform.transform_points(p)
        idx = geom_in.physical_to_index(q)
        inside = _inside(idx, geom_in.size)
        v = _trilinear_clamped(prob, idx[..., ::-1])
        out[lin] = np.where(inside, v, default)
    return out.reshape(D, H, W)


def closed_form_index_map(j_xyz, N_A, N_B, n):
    """SURVEY App. C.1 helper for the property test: q = (j+1/2) n/N_B - 1/2 ; i = (q'+1/2) N_A/n - 1/2."""
    j = np.asarray(j_xyz, dtype=np.float64)
    return (j + 0.5) * np.asarray(n) / np.asarray(N_B) - 0.5
