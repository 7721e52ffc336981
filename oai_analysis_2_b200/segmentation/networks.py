"""B200 implementation of the reference segmentation UNet (oai_analysis/segmentation/networks.py:38-149).

`UNet` keeps the reference constructor, state_dict keys and `weights_init` semantics, but is not an nn.Module: its
forward is a sequence of C-ABI kernel launches (tcgen05 implicit-GEMM convolutions + stem / pool / head kernels)
over channels-last 16-bit activations.  BatchNorm (eval) is folded into the preceding convolution, transposed
convolutions are re-expressed as convolutions / pointwise GEMMs, and the skip concatenations are two-source K loops.
"""
import math
import os

import numpy as np
import torch

from .. import ops

# name, kind ("c" Conv3d k3 p1 | "t3" ConvTranspose3d k3 s1 p1 | "t2" ConvTranspose3d k2 s2), cin, cout
_LAYERS = (
    ("ec0", "c", None, 32), ("ec1", "c", 32, 64), ("ec2", "c", 64, 64), ("ec3", "c", 64, 128),
    ("ec4", "c", 128, 128), ("ec5", "c", 128, 256), ("ec6", "c", 256, 256), ("ec7", "c", 256, 512),
    ("dc9", "t2", 512, 512), ("dc8", "t3", 768, 256), ("dc7", "t3", 256, 256), ("dc6", "t2", 256, 256),
    ("dc5", "t3", 384, 128), ("dc4", "t3", 128, 128), ("dc3", "t2", 128, 128), ("dc2", "t3", 192, 64),
    ("dc1", "t3", 64, 64),
)
# decoder layers whose input is cat((upsampled, skip), dim=1) (networks.py:127,134,141): channels of the first source
_SPLIT = {"dc8": 512, "dc5": 256, "dc2": 128}


class UNet:
    """Drop-in for networks.UNet(in_channels, n_classes, bias=False, BN=False) on the prediction path."""

    def __init__(self, in_channels, n_classes, bias=False, BN=False):
        if in_channels != 1:
            raise NotImplementedError("the fused stem kernel supports in_channels == 1 (the OAI DESS configuration)")
        self.in_channel = in_channels
        self.n_classes = n_classes
        self.bias = bias
        self.BN = BN
        self.device = torch.device("cpu")
        self.precision = "fp16"  # fp16 carries TF32's 10-bit mantissa: the reference's own cuDNN default precision
        # ConvTranspose3d(k2,s2): one stacked-tap launch (default; reads the input once) or 8 pointwise launches
        self.up2_single_launch = os.environ.get("OAI_B200_UP2_SINGLE", "1") == "1"
        self._sd = self._blank_state_dict()
        self._packed = {}

    # ------------------------------------------------------------------ nn.Module-like surface used by the reference
    def layer_table(self):
        return [(n, k, self.in_channel if ci is None else ci, co) for n, k, ci, co in _LAYERS]

    def _blank_state_dict(self):
        sd = {}
        for name, kind, ci, co in self.layer_table():
            k = 2 if kind == "t2" else 3
            sd[f"{name}.0.weight"] = torch.zeros((co, ci, k, k, k) if kind == "c" else (ci, co, k, k, k))
            if self.bias:
                sd[f"{name}.0.bias"] = torch.zeros(co)
            if self.BN:
                sd[f"{name}.1.weight"] = torch.ones(co)
                sd[f"{name}.1.bias"] = torch.zeros(co)
                sd[f"{name}.1.running_mean"] = torch.zeros(co)
                sd[f"{name}.1.running_var"] = torch.ones(co)
                sd[f"{name}.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        sd["dc0.weight"] = torch.zeros(self.n_classes, 64, 1, 1, 1)
        if self.bias:
            sd["dc0.bias"] = torch.zeros(self.n_classes)
        return sd

    def weights_init(self):
        """networks.py:71-78: xavier_normal_ on every conv weight, zero bias, BatchNorm left at its defaults."""
        for key, w in self._sd.items():
            if key.endswith(".0.weight") or key == "dc0.weight":
                torch.nn.init.xavier_normal_(w)
            elif key.endswith(".0.bias") or key == "dc0.bias":
                w.zero_()
        self._packed = {}

    def state_dict(self):
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict=True):
        missing = [k for k in self._sd if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for UNet: missing {missing}, unexpected {unexpected}")
        for k in self._sd:
            if k in state_dict:
                v = torch.as_tensor(state_dict[k]).detach().cpu()
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._sd[k].shape)}")
                self._sd[k] = v.to(self._sd[k].dtype).clone()
        self._packed = {}

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("oai_analysis_2_b200 UNet runs on CUDA (sm_100a) only; got device %r" % (device,))
        self.device = device
        self._packed = {}
        return self

    def cuda(self):
        return self.to("cuda")

    def eval(self):
        return self

    # ------------------------------------------------------------------ weight re-packing
    def _folded(self, name, kind):
        """Conv weight in Conv3d orientation [co, ci, k,k,k] (float64) and bias with BatchNorm(eval) folded in."""
        w = self._sd[f"{name}.0.weight"].double()
        if kind != "c":
            w = w.transpose(0, 1)  # ConvTranspose stores [ci, co, ...]
            if kind == "t3":
                w = w.flip(2, 3, 4)  # stride-1 transposed conv == conv with the spatially flipped filter
        co = w.shape[0]
        b = self._sd[f"{name}.0.bias"].double() if self.bias else torch.zeros(co, dtype=torch.float64)
        if self.BN:
            g, beta = self._sd[f"{name}.1.weight"].double(), self._sd[f"{name}.1.bias"].double()
            mean, var = self._sd[f"{name}.1.running_mean"].double(), self._sd[f"{name}.1.running_var"].double()
            s = g / torch.sqrt(var + 1e-5)
            w = w * s.view(-1, 1, 1, 1, 1)
            b = (b - mean) * s + beta
        return w.contiguous(), b

    def _fmt(self):
        return {"fp16": 0, "bf16": 1}[self.precision]

    def prepare(self, tile_zyx):
        """Pack every layer for tiles of the given (z, y, x) size; cached per tile size."""
        key = (tuple(int(v) for v in tile_zyx), self.precision)
        if key in self._packed:
            return self._packed[key]
        td, th, tw = key[0]
        if td % 8 or th % 8 or tw % 8:
            raise ValueError(f"tile size {key[0]} must be divisible by 8 (three 2x poolings)")
        fmt, dev = self._fmt(), self.device
        P = {}
        level = {"ec0": 0, "ec1": 0, "ec2": 1, "ec3": 1, "ec4": 2, "ec5": 2, "ec6": 3, "ec7": 3, "dc9": 3, "dc8": 2,
                 "dc7": 2, "dc6": 2, "dc5": 1, "dc4": 1, "dc3": 1, "dc2": 0, "dc1": 0}
        for name, kind, ci, co in self.layer_table():
            w, b = self._folded(name, kind)
            D, H, W = (td >> level[name], th >> level[name], tw >> level[name])
            bias = b.float().to(dev)
            if name == "ec0":
                P[name] = dict(w=w.float().reshape(co, 27).t().contiguous().to(dev), b=bias, cout=co)
            elif kind == "t2":
                P[name] = dict(b=bias, cout=co, dims=(D, H, W))
                if self.up2_single_launch:
                    P[name]["w"] = ops.pack_convt2_weights(w.float(), D, H, W, fmt, device=dev)
                else:
                    P[name]["taps"] = [ops.pack_conv_weights(w[:, :, a, bb, c].float(), ci, 0, D, H, W, True, fmt,
                                                             device=dev)
                                       for a in range(2) for bb in range(2) for c in range(2)]
            else:
                c0 = _SPLIT.get(name, ci)
                P[name] = dict(w=ops.pack_conv_weights(w.float(), c0, ci - c0, D, H, W, False, fmt, device=dev),
                               b=bias, cout=co, dims=(D, H, W))
        P["dc0"] = dict(w=self._sd["dc0.weight"].float().reshape(self.n_classes, 64).contiguous().to(dev),
                        b=(self._sd["dc0.bias"].float() if self.bias else torch.zeros(self.n_classes)).to(dev))
        self._packed[key] = P
        return P

    # ------------------------------------------------------------------ dead-halo regions
    @staticmethod
    def needed_regions(tile_zyx, overlap_zyx):
        """Output sub-boxes (inclusive lo/hi per z,y,x) the kept tile interior actually depends on, per decoder
        layer.  The reference computes every layer on the whole tile and crops afterwards
        (image_transforms.py:497-503); a k3 conv widens the needed box by 1, a k2s2 up-conv halves it."""
        t, o = np.asarray(tile_zyx), np.asarray(overlap_zyx)

        def box(lo, hi, lvl):
            return np.maximum(lo, 0), np.minimum(hi, (t >> lvl) - 1)

        def dil(b, k, lvl):
            return box(b[0] - k, b[1] + k, lvl)

        R = {}
        n0 = box(o, t - o - 1, 0)
        R["dc1"], R["dc2"] = n0, dil(n0, 1, 0)
        need = dil(n0, 2, 0)
        for up, a, b, lvl in (("dc3", "dc4", "dc5", 1), ("dc6", "dc7", "dc8", 2)):
            q = box(need[0] // 2, need[1] // 2, lvl)
            R[up], R[a], R[b] = q, q, dil(q, 1, lvl)
            need = dil(q, 2, lvl)
        R["dc9"] = box(need[0] // 2, need[1] // 2, 3)
        return R

    @staticmethod
    def _region_arg(box, dims, cout, pointwise=False):
        """(d_lo, d_cnt, h_lo, h_cnt) for the C ABI (the kernel groups d-slices by itself; the last group may be
        partial)."""
        lo, hi = box
        return (int(lo[0]), int(hi[0] - lo[0] + 1), int(lo[1]), int(hi[1] - lo[1] + 1))

    # ------------------------------------------------------------------ forward pieces
    def _conv(self, P, name, src0, src1=None, box=None):
        L = P[name]
        region = None if box is None else self._region_arg(box, L["dims"], L["cout"])
        return ops.conv3d_igemm(src0, src1, L["w"], L["b"], L["cout"], False, True, self._fmt(), region=region)

    def _up(self, P, name, src, box=None):
        """ConvTranspose3d(k=2, s=2) + ReLU: one launch, the 8 sub-filters stacked along N, scattered into the 2x grid."""
        L = P[name]
        co = L["cout"]
        region = None if box is None else self._region_arg(box, L["dims"], co, True)
        if "w" in L:
            return ops.convt2_igemm(src, L["w"], L["b"], co, True, self._fmt(), region)
        NT, D, H, W, _ = src.shape
        out = torch.empty((NT, 2 * D, 2 * H, 2 * W, co), dtype=src.dtype, device=src.device)
        sW, sH, sD, sN = 2 * co, 4 * W * co, 8 * H * W * co, 8 * D * H * W * co
        for t in range(8):
            a, b, c = t >> 2, (t >> 1) & 1, t & 1
            ops.conv3d_igemm(src, None, L["taps"][t], L["b"], co, True, True, self._fmt(), out=out,
                             out_view=(((a * 2 * H + b) * 2 * W + c) * co, sN, sD, sH, sW), region=region)
        return out

    def forward_features(self, P, e0, overlap_zyx=(0, 0, 0)):
        """networks.py:110-144 from ec1 to dc2 on act16 tensors; e0 is the stem (ec0) output.  With a non-zero
        overlap only the part of each decoder layer the kept interior depends on is computed."""
        fmt = self._fmt()
        B = self.needed_regions(e0.shape[1:4], overlap_zyx) if any(overlap_zyx) else {}
        B = {k: B.get(k) for k in ("dc2", "dc3", "dc4", "dc5", "dc6", "dc7", "dc8", "dc9")}
        syn0 = self._conv(P, "ec1", e0)
        del e0
        syn1 = self._conv(P, "ec3", self._conv(P, "ec2", ops.maxpool2(syn0, fmt)))
        syn2 = self._conv(P, "ec5", self._conv(P, "ec4", ops.maxpool2(syn1, fmt)))
        e7 = self._conv(P, "ec7", self._conv(P, "ec6", ops.maxpool2(syn2, fmt)))
        d7 = self._conv(P, "dc7", self._conv(P, "dc8", self._up(P, "dc9", e7, B["dc9"]), syn2, B["dc8"]), None,
                        B["dc7"])
        del e7, syn2
        d4 = self._conv(P, "dc4", self._conv(P, "dc5", self._up(P, "dc6", d7, B["dc6"]), syn1, B["dc5"]), None,
                        B["dc4"])
        del d7, syn1
        d2 = self._conv(P, "dc2", self._up(P, "dc3", d4, B["dc3"]), syn0, B["dc2"])
        del d4, syn0
        return d2

    def head(self, P, d2, out, geom, tile0, crop_zyx, out_mode):
        """dc1 + dc0 + sigmoid + assemble in one launch (networks.py:145-148, segmenter.py:121-129)."""
        L = P["dc1"]
        return ops.conv3d_igemm_head(d2, L["w"], L["b"], P["dc0"]["w"], P["dc0"]["b"], out, geom, tile0, crop_zyx,
                                     out_mode, self._fmt())

    def forward(self, x):
        """Module-style forward on explicit tiles: x [N, 1, D, H, W] float32 (cuda) -> logits [N, n_classes, D, H, W].

        Provided for API parity with the reference model object; Segmenter3DInPatchClassWise uses the fused
        volume path (stem reads the volume in place, head writes the assembled maps)."""
        N, _, td, th, tw = x.shape
        P = self.prepare((td, th, tw))
        # a stack of tiles is a "volume" whose tiling has zero overlap and a (N,1,1) grid
        vol = x.reshape(N * td, th, tw).contiguous().float()
        geom = ops.make_geom((td, th, tw), (td, th, tw), (0, 0, 0), (N, 1, 1))
        e0 = ops.seg_stem(vol, geom, 0, N, P["ec0"]["w"], P["ec0"]["b"], self._fmt())
        d2 = self.forward_features(P, e0)
        out = torch.empty((self.n_classes, N * td, th, tw), dtype=torch.float32, device=x.device)
        self.head(P, d2, out, geom, 0, (0, 0, 0), 2)
        return out.view(self.n_classes, N, td, th, tw).transpose(0, 1)

    __call__ = forward


network_dic = {"UNet": UNet}


def get_network(name):
    """networks.py:858-866 (the reference returns None for unknown names; here that is an error)."""
    if name not in network_dic:
        raise KeyError(f"Network {name} is not implemented in the B200 path (available: {sorted(network_dic)})")
    return network_dic[name]
