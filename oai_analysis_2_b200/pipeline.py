"""The full per-knee hot path on one GPU (BASELINE config 3): segmentation -> GradICON registration to the atlas ->
warp of the FC/TC probability maps onto the atlas grid -> warp of thickness-mesh vertices into atlas space.

This is what notebooks/FullDemo.ipynb cells 4-7 and the Dask chain of dask_processing.py (segment_method,
register_images_delayed, deform_probmap_delayed x2) do per knee, with every intermediate kept in HBM: one H2D of the
input volume, one D2H of the atlas-space maps, the displacement fields and the warped vertices."""
import numpy as np
import torch

from . import ops
from .icon_registration import itk_wrapper
from .transforms import CompositeTransform, Geometry


class KneePipeline:
    def __init__(self, segmenter, reg_model, atlas_array, atlas_geom=None, device="cuda"):
        self.device = torch.device(device)
        self.segmenter = segmenter
        self.reg_model = reg_model.to(self.device)
        atlas_array = np.ascontiguousarray(atlas_array, dtype=np.float32)
        self.atlas = torch.from_numpy(atlas_array).to(self.device)
        self.atlas_geom = atlas_geom or Geometry(atlas_array.shape[::-1])
        self._pinned = {}

    def run_device(self, vol, geom, vertices=None):
        """vol: float32 [D,H,W] on the device (intensities windowed to [0,1]); vertices: float64 [n,3] physical
        points in the knee's space.  Returns device tensors."""
        prob = self.segmenter.segment_device(vol, if_output_prob_map=True,
                                             tiles_per_batch=self.segmenter.config.get("tiles_per_batch"))
        phi_AB, phi_BA = itk_wrapper.register_pair_device(self.reg_model, vol, self.atlas)
        tr_AB = CompositeTransform(ops.displacement_field(phi_AB[0]), geom, self.atlas_geom)
        tr_BA = CompositeTransform(ops.displacement_field(phi_BA[0]), self.atlas_geom, geom)
        warped = tr_AB.resample_device(prob, geom, self.atlas_geom)      # FC, TC on the atlas grid
        out = dict(prob=prob, warped=warped, phi_AB=tr_AB, phi_BA=tr_BA)
        if vertices is not None:
            # ITK resampling transforms map output-space points to input space, so patient -> atlas is phi_BA
            out["vertices"] = ops.warp_points(vertices, tr_BA.disp, tr_BA.from_network_space_inv,
                                              tr_BA.to_network_space)
        return out

    def _pin(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in self._pinned:
            self._pinned[key] = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        return self._pinned[key]

    def run(self, volume, geom=None, vertices=None, return_fields=True):
        """Host in, host out (the e2e path): `volume` float32 [D,H,W] numpy / pinned tensor; returns numpy arrays
        (views of reused pinned buffers) + the two transforms."""
        vol_h = torch.as_tensor(volume)
        geom = geom or Geometry(tuple(vol_h.shape)[::-1])
        vol = vol_h.to(self.device, non_blocking=True)
        verts = None
        if vertices is not None:
            verts = torch.as_tensor(vertices, dtype=torch.float64).to(self.device, non_blocking=True)
        r = self.run_device(vol, geom, verts)
        res = {}
        w = self._pin("warped", r["warped"].shape, torch.float32)
        w.copy_(r["warped"], non_blocking=True)
        res["FC_atlas"], res["TC_atlas"] = w[0].numpy(), w[1].numpy()
        if return_fields:
            for k in ("phi_AB", "phi_BA"):
                f = self._pin(k, r[k].disp.shape, torch.float32)
                f.copy_(r[k].disp, non_blocking=True)
                res[k + "_field"] = f.numpy()
        if verts is not None:
            v = self._pin("verts", r["vertices"].shape, torch.float64)
            v.copy_(r["vertices"], non_blocking=True)
            res["vertices_atlas"] = v.numpy()
        torch.cuda.current_stream().synchronize()
        res["phi_AB"], res["phi_BA"] = r["phi_AB"], r["phi_BA"]
        res["d2h_bytes"] = sum(t.numel() * t.element_size() for (n, _, _), t in self._pinned.items()
                               if n in ("warped", "verts") or (return_fields and n.startswith("phi")))
        res["h2d_bytes"] = vol_h.numel() * 4 + (0 if verts is None else verts.numel() * 8)
        return res
