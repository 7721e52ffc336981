"""The full per-knee hot path on one GPU (BASELINE config 3): segmentation -> GradICON registration to the atlas ->
warp of the FC/TC probability maps onto the atlas grid -> warp of thickness-mesh vertices into atlas space.

This is what notebooks/FullDemo.ipynb cells 4-7 and the Dask chain of dask_processing.py (segment_method,
register_images_delayed, deform_probmap_delayed x2) do per knee, with every intermediate kept in HBM: one H2D of the
input volume, one D2H of the atlas-space maps, the displacement fields and the warped vertices."""
import numpy as np
import torch

from . import ops
from .icon_registration import itk_wrapper
from .transforms import CompositeTransform, Geometry, HostFieldTransform


class KneePipeline:
    def __init__(self, segmenter, reg_model, atlas_array, atlas_geom=None, device="cuda"):
        self.device = torch.device(device)
        self.segmenter = segmenter
        self.reg_model = reg_model.to(self.device)
        atlas_array = np.ascontiguousarray(atlas_array, dtype=np.float32)
        self.atlas = torch.from_numpy(atlas_array).to(self.device)
        self.atlas_geom = atlas_geom or Geometry(atlas_array.shape[::-1])
        self._pinned = {}

    def run_device(self, vol, geom, vertices=None, overlap_registration=False):
        """vol: float32 [D,H,W] on the device (intensities windowed to [0,1]); vertices: float64 [n,3] physical
        points in the knee's space.  Returns device tensors.

        overlap_registration: registration needs only the volume and the atlas, so it can run as a parallel branch
        beside the segmentation (a second stream; inside capture() a second branch of the graph).  The persistent conv
        kernel owns every SM's register file while it runs, so the branch only fills the HBM-bound phases (stem, pools)
        and the tails of the conv launches.  Used by capture() only: in eager mode the caching allocator would need
        record_stream bookkeeping for the tensors that cross streams."""
        nvtx = torch.cuda.nvtx   # one range per stage (the reference has no tracing at all)

        def register():
            nvtx.range_push("oai.registration")
            phi_AB, phi_BA = itk_wrapper.register_pair_device(self.reg_model, vol, self.atlas)
            a = CompositeTransform(ops.displacement_field(phi_AB[0]), geom, self.atlas_geom)
            b = CompositeTransform(ops.displacement_field(phi_BA[0]), self.atlas_geom, geom)
            nvtx.range_pop()
            return a, b

        def segment():
            nvtx.range_push("oai.segmentation")
            p = self.segmenter.segment_device(vol, if_output_prob_map=True,
                                              tiles_per_batch=self.segmenter.config.get("tiles_per_batch"))
            nvtx.range_pop()
            return p

        if overlap_registration:
            cur = torch.cuda.current_stream(self.device)
            if not hasattr(self, "_reg_stream"):
                self._reg_stream = torch.cuda.Stream(device=self.device)
            self._reg_stream.wait_stream(cur)
            with torch.cuda.stream(self._reg_stream):
                tr_AB, tr_BA = register()
            prob = segment()
            cur.wait_stream(self._reg_stream)
        else:
            prob = segment()
            tr_AB, tr_BA = register()
        nvtx.range_push("oai.warp_probmaps")
        warped = tr_AB.resample_device(prob, geom, self.atlas_geom)      # FC, TC on the atlas grid
        nvtx.range_pop()
        out = dict(prob=prob, warped=warped, phi_AB=tr_AB, phi_BA=tr_BA)
        if vertices is not None:
            # ITK resampling transforms map output-space points to input space, so patient -> atlas is phi_BA
            nvtx.range_push("oai.warp_vertices")
            out["vertices"] = ops.warp_points(vertices, tr_BA.disp, tr_BA.from_network_space_inv,
                                              tr_BA.to_network_space)
            nvtx.range_pop()
        return out

    # -- CUDA graph of the whole per-knee path (about a hundred launches; replaying one graph removes the launch gaps)
    def capture(self, vol_shape, geom, n_vertices=None, on_record=None, overlap_registration=False):
        """Record run_device for volumes of `vol_shape` (and `n_vertices` mesh vertices) into a CUDA graph with static
        input / output buffers.  Afterwards run_device_graph / run replay it.  The graph holds the addresses of the
        models' cached buffers (packed weights, the registration nets' concatenation buffers), so after capture this
        segmenter / registration model must not be run eagerly on other shapes (that would re-size those caches)."""
        dev = self.device
        self._g_vol = torch.zeros(tuple(vol_shape), dtype=torch.float32, device=dev)
        self._g_verts = None if not n_vertices else torch.zeros((int(n_vertices), 3), dtype=torch.float64, device=dev)
        self._g_geom = geom
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # eager pass: packs weights, sizes the kernels' static buffers
            self.run_device(self._g_vol, geom, self._g_verts)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()
        self._graph = torch.cuda.CUDAGraph()
        if on_record is not None:
            on_record()   # e.g. switch the library's per-launch profiling events on for the recorded pass only
        with torch.cuda.graph(self._graph):
            self._g_out = self.run_device(self._g_vol, geom, self._g_verts, overlap_registration=overlap_registration)
        return self

    def release_graph(self):
        """Drop the captured graph, its static buffers and the stream API's staging buffers (tens of GB of private
        pool memory go back to the allocator)."""
        for name in ("_graph", "_g_out", "_g_vol", "_g_verts", "_rs"):
            if hasattr(self, name):
                delattr(self, name)
        torch.cuda.synchronize(self.device)
        torch.cuda.empty_cache()

    def _graph_matches(self, vol, geom, vertices):
        return (getattr(self, "_graph", None) is not None and tuple(vol.shape) == tuple(self._g_vol.shape)
                and geom is self._g_geom
                and ((vertices is None) == (self._g_verts is None))
                and (vertices is None or tuple(vertices.shape) == tuple(self._g_verts.shape)))

    def run_device_graph(self, vol, geom, vertices=None):
        """Same contract as run_device; the returned tensors AND transforms alias the graph's static outputs: they are
        overwritten by the next replay (clone what must outlive it; run() / run_stream() hand out owned copies)."""
        if not self._graph_matches(vol, geom, vertices):
            return self.run_device(vol, geom, vertices)
        self._g_vol.copy_(vol, non_blocking=True)
        if vertices is not None:
            self._g_verts.copy_(vertices, non_blocking=True)
        self._graph.replay()
        return self._g_out

    def run_stream(self, items, return_fields=True):
        """Throughput form of run() for a sequence of knees: `items` yields (volume, vertices) host buffers (pinned for
        real overlap) matching the captured graph.  The H2D copy of knee i+1 and the D2H copy of knee i-1 run on their
        own streams while knee i computes (device staging buffers on both sides of the graph's static I/O).  Yields one
        result dict per knee, in order; its arrays are views of double-buffered pinned memory and stay valid until the
        generator is advanced again (copy what must outlive that)."""
        if getattr(self, "_graph", None) is None:
            raise RuntimeError("run_stream needs KneePipeline.capture() first")
        dev = self.device
        comp = torch.cuda.current_stream(dev)
        if not hasattr(self, "_rs"):
            g = self._g_out
            outs = {"warped": g["warped"], "phi_AB": g["phi_AB"].disp, "phi_BA": g["phi_BA"].disp}
            if self._g_verts is not None:
                outs["verts"] = g["vertices"]
            self._rs = dict(
                h2d=torch.cuda.Stream(device=dev), d2h=torch.cuda.Stream(device=dev), outs=outs,
                vol=[torch.empty_like(self._g_vol) for _ in range(2)],
                verts=[None if self._g_verts is None else torch.empty_like(self._g_verts) for _ in range(2)],
                dstage=[{k: torch.empty_like(v) for k, v in outs.items()} for _ in range(2)],
                host=[{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in outs.items()}
                      for _ in range(2)],
                ev_h2d=[torch.cuda.Event() for _ in range(2)], ev_comp=[torch.cuda.Event() for _ in range(2)],
                ev_d2h=[torch.cuda.Event() for _ in range(2)])
        rs = self._rs
        keys = [k for k in rs["outs"] if return_fields or not k.startswith("phi")]

        def upload(item, b):
            vol_h, verts_h = item
            with torch.cuda.stream(rs["h2d"]):
                rs["h2d"].wait_event(rs["ev_comp"][b])      # the compute that last read this staging slot is done
                rs["vol"][b].copy_(torch.as_tensor(vol_h), non_blocking=True)
                if rs["verts"][b] is not None:
                    rs["verts"][b].copy_(torch.as_tensor(verts_h, dtype=torch.float64), non_blocking=True)
                rs["ev_h2d"][b].record(rs["h2d"])

        def result(b):
            rs["ev_d2h"][b].synchronize()
            h = rs["host"][b]
            res = {"FC_atlas": h["warped"][0].numpy(), "TC_atlas": h["warped"][1].numpy()}
            if return_fields:
                res["phi_AB_field"], res["phi_BA_field"] = h["phi_AB"].numpy(), h["phi_BA"].numpy()
                # this knee's own transforms: the graph's static displacement buffers already hold a later knee, so the
                # transform objects are rebuilt around the (pinned) host copy of THIS knee's fields
                g = self._g_out
                res["phi_AB"] = HostFieldTransform(h["phi_AB"], g["phi_AB"].geom_A, g["phi_AB"].geom_B, self.device)
                res["phi_BA"] = HostFieldTransform(h["phi_BA"], g["phi_BA"].geom_A, g["phi_BA"].geom_B, self.device)
            if "verts" in h:
                res["vertices_atlas"] = h["verts"].numpy()
            res["d2h_bytes"] = sum(h[k].numel() * h[k].element_size() for k in keys)
            res["h2d_bytes"] = self._g_vol.numel() * 4 + (0 if self._g_verts is None else self._g_verts.numel() * 8)
            return res

        it = iter(items)
        nxt = next(it, None)
        if nxt is None:
            return
        upload(nxt, 0)
        i = 0
        while nxt is not None:
            b = i & 1
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, b ^ 1)                           # flies while knee i computes
            comp.wait_event(rs["ev_h2d"][b])
            self._g_vol.copy_(rs["vol"][b], non_blocking=True)
            if self._g_verts is not None:
                self._g_verts.copy_(rs["verts"][b], non_blocking=True)
            self._graph.replay()
            comp.wait_event(rs["ev_d2h"][b])                 # the D2H that last read this output slot is done
            for k in keys:
                rs["dstage"][b][k].copy_(rs["outs"][k], non_blocking=True)
            rs["ev_comp"][b].record(comp)
            with torch.cuda.stream(rs["d2h"]):
                rs["d2h"].wait_event(rs["ev_comp"][b])
                for k in keys:
                    rs["host"][b][k].copy_(rs["dstage"][b][k], non_blocking=True)
                rs["ev_d2h"][b].record(rs["d2h"])
            if i > 0:
                yield result(b ^ 1)
            i += 1
        yield result((i - 1) & 1)

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
        verts_h = None if vertices is None else torch.as_tensor(vertices, dtype=torch.float64)
        if self._graph_matches(vol_h, geom, verts_h):
            r = self.run_device_graph(vol_h, geom, verts_h)   # H2D straight into the graph's static inputs
            verts = verts_h
        else:
            vol = vol_h.to(self.device, non_blocking=True)
            verts = None if verts_h is None else verts_h.to(self.device, non_blocking=True)
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
        if getattr(self, "_graph", None) is not None and r is self._g_out:
            # graph path: the static displacement buffers are overwritten by the next replay -- hand out transforms
            # that own a copy of this knee's fields
            res["phi_AB"] = CompositeTransform(r["phi_AB"].disp.clone(), r["phi_AB"].geom_A, r["phi_AB"].geom_B)
            res["phi_BA"] = CompositeTransform(r["phi_BA"].disp.clone(), r["phi_BA"].geom_A, r["phi_BA"].geom_B)
        else:
            res["phi_AB"], res["phi_BA"] = r["phi_AB"], r["phi_BA"]
        res["d2h_bytes"] = sum(t.numel() * t.element_size() for (n, _, _), t in self._pinned.items()
                               if n in ("warped", "verts") or (return_fields and n.startswith("phi")))
        res["h2d_bytes"] = vol_h.numel() * 4 + (0 if verts is None else verts.numel() * 8)
        return res
