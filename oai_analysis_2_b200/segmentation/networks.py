"""B200 implementation of the reference segmentation UNet (oai_analysis/segmentation/networks.py:38-149).

`UNet` keeps the reference constructor, state_dict keys and `weights_init` semantics, but is not an nn.Module: it
holds the state dict and hands it to the library's stage-level entry points (oai_seg_create / oai_seg_forward,
csrc/seg_net.cu), where BatchNorm (eval) is folded into the preceding convolution, transposed convolutions are
re-expressed as convolutions / pointwise GEMMs, the skip concatenations become two-source K loops and the layers run as
tcgen05 implicit-GEMM launches over channels-last 16-bit activations.
"""
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
        # "mixed" (default) meets the north-star Dice bar: fp16 operands (TF32's mantissa, the reference's own cuDNN
        # default) except that the two full-resolution decoder layers read fp16 hi+lo activations (dc2: its skip input; dc1).  "fp16" is the fast
        # all-16-bit plan, "fp16x2" / "fp16x3" split every layer (x3 is fp32-faithful), "bf16" for range over precision.
        self.precision = os.environ.get("OAI_B200_SEG_PRECISION", "mixed")
        self._sd = self._blank_state_dict()
        self._handles = {}

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
        self._handles = {}

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
        self._handles = {}

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("oai_analysis_2_b200 UNet runs on CUDA (sm_100a) only; got device %r" % (device,))
        self.device = device
        self._handles = {}
        return self

    def cuda(self):
        return self.to("cuda")

    def eval(self):
        return self

    # ------------------------------------------------------------------ stage-level C ABI
    PRECISIONS = {"fp16": (0, 0), "bf16": (0, 1), "mixed": (1, 0), "fp16x2": (2, 0), "fp16x3": (3, 0)}
    LAYER_NAMES = tuple(n for n, _, _, _ in _LAYERS)

    def _fmt(self):
        return self.PRECISIONS[self.precision][1]

    def layer_terms(self):
        """Products K-concatenated per layer (ec0..dc1): 1 = 16-bit operands, 2 = fp16 hi+lo activations, 3 = both."""
        return ops.seg_layer_terms(self.PRECISIONS[self.precision][0])

    def seg_handle(self, patch_xyz, overlap_xyz):
        """oai_seg_create for this model's weights: BatchNorm folding, convT re-orientation, weight packing, dead-halo
        regions and the layer sequence all live behind the handle (csrc/seg_net.cu).  Cached per geometry."""
        if self.device.type != "cuda":
            raise RuntimeError("oai_analysis_2_b200 UNet runs on CUDA (sm_100a) only; call .to('cuda') first")
        key = (tuple(int(v) for v in patch_xyz), tuple(int(v) for v in overlap_xyz), self.precision, self.device)
        if key not in self._handles:
            plan, fmt = self.PRECISIONS[self.precision]
            with torch.cuda.device(self.device):
                self._handles[key] = ops.SegHandle(self._sd, self.in_channel, self.n_classes, self.bias, self.BN,
                                                   key[0], key[1], fmt, plan)
        return self._handles[key]

    @staticmethod
    def needed_regions(tile_zyx, overlap_zyx):
        """Output sub-boxes (inclusive lo/hi per z,y,x) the kept tile interior depends on, per decoder layer: the
        library's own table (oai_seg_needed_regions), keyed by layer name."""
        boxes = ops.seg_needed_regions(tile_zyx, overlap_zyx)
        return {n: (boxes[i, :3].copy(), boxes[i, 3:].copy()) for i, n in enumerate(UNet.LAYER_NAMES) if n.startswith("dc")}

    @staticmethod
    def _region_arg(box, dims=None, cout=None, pointwise=False):
        """(d_lo, d_cnt, h_lo, h_cnt) as the conv C ABI takes it (the kernel groups d-slices by itself)."""
        lo, hi = box
        return (int(lo[0]), int(hi[0] - lo[0] + 1), int(lo[1]), int(hi[1] - lo[1] + 1))

    def forward(self, x):
        """Module-style forward on explicit tiles: x [N, 1, D, H, W] float32 (cuda) -> logits [N, n_classes, D, H, W].

        Provided for API parity with the reference model object; Segmenter3DInPatchClassWise uses the volume path
        (the stem reads the volume in place, the head writes the assembled maps)."""
        N, _, td, th, tw = x.shape
        # a stack of tiles is a "volume" whose tiling has zero overlap and an (N,1,1) grid
        h = self.seg_handle((tw, th, td), (0, 0, 0))
        vol = x.reshape(N * td, th, tw).contiguous().float()
        out = h.forward(vol, out_mode=2)
        return out.view(self.n_classes, N, td, th, tw).transpose(0, 1)

    __call__ = forward


network_dic = {"UNet": UNet}


def get_network(name):
    """networks.py:858-866 (the reference returns None for unknown names; here that is an error)."""
    if name not in network_dic:
        raise KeyError(f"Network {name} is not implemented in the B200 path (available: {sorted(network_dic)})")
    return network_dic[name]
