"""ITK-shaped transform objects produced by registration and consumed by the warps.

`CompositeTransform` is what icon's create_itk_transform builds (itk.CompositeTransform[D,3] holding
to_network_space o DisplacementFieldTransform o from_network_space^-1): T(p) = R_A( D( R_B^-1(p) ) ).  It keeps the
displacement field on the device (float32, [D,H,W,3], x,y,z components in network-voxel units) for the GPU warps and
exposes it as a float64 host array like the ITK image would.
"""
import numpy as np
import torch

from . import itk_compat, ops


class Geometry:
    """size / spacing / origin / direction of an image grid, ITK convention (x,y,z)."""

    def __init__(self, size_xyz, spacing=(1, 1, 1), origin=(0, 0, 0), direction=None):
        self.size = np.asarray(size_xyz, dtype=np.int64)
        self.spacing = np.asarray(spacing, dtype=np.float64)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64)

    @classmethod
    def of(cls, image):
        arr = itk_compat.array_from_image(image)
        sp, org, dr = itk_compat.image_metadata(image)
        return cls(arr.shape[::-1], sp, org, dr)

    def index_to_physical_affine(self):
        return self.direction @ np.diag(self.spacing), self.origin.copy()

    def physical_to_index_affine(self):
        Minv = np.diag(1.0 / self.spacing) @ np.linalg.inv(self.direction)
        return Minv, -Minv @ self.origin


def resampling_transform(geom, net_shape_xyz):
    """icon itk_wrapper.resampling_transform: R(x) = M (x - c_f) + c_m as (M, t) with t = c_m - M c_f."""
    n = np.asarray(net_shape_xyz, dtype=np.float64)
    c_f = (n - 1) / 2.0
    Mi, ti = geom.index_to_physical_affine()
    c_m = Mi @ ((geom.size - 1) / 2.0) + ti
    M = geom.direction @ np.diag(geom.spacing * geom.size / n)
    return M, c_m - M @ c_f


def _compose(a, b):
    """affine a after affine b."""
    return a[0] @ b[0], a[0] @ b[1] + a[1]


def _invert(a):
    Mi = np.linalg.inv(a[0])
    return Mi, -Mi @ a[1]


class CompositeTransform:
    def __init__(self, disp_dev, geom_A, geom_B):
        self.disp = disp_dev                      # cuda float32 [D,H,W,3]
        d, h, w = disp_dev.shape[:3]
        self.net_xyz = np.array([w, h, d])
        self.geom_A, self.geom_B = geom_A, geom_B
        self.to_network_space = resampling_transform(geom_A, self.net_xyz)
        self.from_network_space_inv = _invert(resampling_transform(geom_B, self.net_xyz))

    # ---- ITK-like accessors
    def GetNumberOfTransforms(self):
        return 3

    def displacement_field_array(self):
        """float64 [D,H,W,3] like itk.array_from_image(tr.GetDisplacementField()).  Read from the device every time:
        the field buffer may be a CUDA graph's static output that a later replay rewrites."""
        return self.disp.cpu().to(torch.float64).numpy()

    def transform_points(self, pts_xyz):
        """TransformPoint over an [n,3] array of physical points (float64).  Runs on the GPU."""
        p = torch.as_tensor(np.ascontiguousarray(pts_xyz, dtype=np.float64)).reshape(-1, 3).to(self.disp.device)
        out = ops.warp_points(p.contiguous(), self.disp, self.from_network_space_inv, self.to_network_space)
        return out.cpu().numpy().reshape(np.shape(pts_xyz))

    def TransformPoint(self, p):
        return self.transform_points(np.asarray(p, dtype=np.float64)[None])[0]

    # ---- resampling (what itk.resample_image_filter(prob, transform=self, ...) computes)
    def resample_device(self, src_dev, geom_src, geom_out, default_value=0.0):
        """src_dev: cuda float32 [C, D, H, W] on geom_src's grid.  Returns [C, *geom_out grid] float32."""
        a = _compose(self.from_network_space_inv, geom_out.index_to_physical_affine())
        b = _compose(geom_src.physical_to_index_affine(), self.to_network_space)
        return ops.warp_volume(src_dev, self.disp, a, b, tuple(int(v) for v in geom_out.size[::-1]), default_value)


class HostFieldTransform(CompositeTransform):
    """CompositeTransform whose displacement field lives in (pinned) host memory -- what KneePipeline.run_stream hands
    out per knee.  It wraps THIS knee's field (never the CUDA graph's static buffer, which already holds a later knee)
    and uploads it on first use by a GPU warp.  Like the result arrays of run_stream it is a view of a double-buffered
    pinned slot: valid until the generator is advanced again; snapshot() returns a transform that owns its field."""

    def __init__(self, field_host, geom_A, geom_B, device, copy=False):
        f = field_host.numpy() if hasattr(field_host, "numpy") else np.asarray(field_host)
        self._field = np.array(f, dtype=np.float32) if copy else f
        self._device = torch.device(device)
        self._dev = None
        d, h, w = self._field.shape[:3]
        self.net_xyz = np.array([w, h, d])
        self.geom_A, self.geom_B = geom_A, geom_B
        self.to_network_space = resampling_transform(geom_A, self.net_xyz)
        self.from_network_space_inv = _invert(resampling_transform(geom_B, self.net_xyz))

    @property
    def disp(self):
        if self._dev is None:
            self._dev = torch.from_numpy(self._field).to(self._device)
        return self._dev

    def displacement_field_array(self):
        return self._field.astype(np.float64)

    def snapshot(self):
        return HostFieldTransform(self._field, self.geom_A, self.geom_B, self._device, copy=True)
