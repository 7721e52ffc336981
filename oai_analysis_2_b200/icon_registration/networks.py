"""tallUNet2 (icon_registration/networks.py::UNet2 with num_layers=5,
channels=[[2,16,32,64,256,512],[16,32,64,128,256]]) as a sequence of fused fp32 CUDA kernels.

The concatenation buffers of the up path are allocated once; the down path writes each level's input directly into
the skip half of its buffer and the up path writes the other half, so torch.cat never runs."""
import numpy as np
import torch

from .. import ops

DOWN = [2, 16, 32, 64, 256, 512]
UP_OUT = [16, 32, 64, 128, 256]
UP_IN = [DOWN[d + 1] + (UP_OUT[d + 1] if d + 1 < 5 else 0) for d in range(5)]


def state_dict_template():
    sd = {}
    for d in range(5):
        sd[f"downConvs.{d}.weight"] = (DOWN[d + 1], DOWN[d], 3, 3, 3)
        sd[f"downConvs.{d}.bias"] = (DOWN[d + 1],)
        sd[f"upConvs.{d}.weight"] = (UP_IN[d], UP_OUT[d], 4, 4, 4)
        sd[f"upConvs.{d}.bias"] = (UP_OUT[d],)
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[f"batchNorms.{d}.{k}"] = (UP_OUT[d],)
    sd["lastConv.weight"] = (3, 18, 3, 3, 3)
    sd["lastConv.bias"] = (3,)
    return sd


class TallUNet2:
    """networks.tallUNet2(dimension=3).  forward(A, B) -> displacement [N,3,D,H,W] (already divided by 10)."""

    def __init__(self):
        # icon initialises lastConv to zero (zero displacement) and everything else with torch defaults
        self._sd = {}
        g = torch.Generator().manual_seed(0)
        tmpl = state_dict_template()
        for k, shape in tmpl.items():
            if k.startswith("lastConv"):
                self._sd[k] = torch.zeros(shape)
            elif k.startswith("batchNorms"):
                self._sd[k] = torch.ones(shape) if k.endswith(("running_var", ".weight")) else torch.zeros(shape)
            else:
                wshape = tmpl[k.rsplit(".", 1)[0] + ".weight"]
                bound = 1.0 / np.sqrt(float(np.prod(wshape[1:])))
                self._sd[k] = torch.empty(shape).uniform_(-bound, bound, generator=g)
        self.device = torch.device("cpu")
        self._packed = None
        self._bufs = {}
        self.use_mma = True  # False: every layer on the fp32 CUDA-core kernels (A/B switch for tests / profiling)

    def state_dict(self, prefix=""):
        return {prefix + k: v for k, v in self._sd.items()}

    def load_state_dict(self, sd, strict=True):
        tmpl = state_dict_template()
        missing = [k for k in tmpl if k not in sd]
        unexpected = [k for k in sd if k not in tmpl and not k.endswith("num_batches_tracked")]
        if strict and (missing or unexpected):
            raise RuntimeError(f"tallUNet2: missing keys {missing}, unexpected keys {unexpected}")
        for k, shape in tmpl.items():
            if k in sd:
                v = torch.as_tensor(sd[k]).detach().float().cpu()
                if tuple(v.shape) != tuple(shape):
                    raise RuntimeError(f"tallUNet2: size mismatch for {k}: {tuple(v.shape)} vs {shape}")
                self._sd[k] = v.clone()
        self._packed = None

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("oai_analysis_2_b200 registration runs on CUDA (sm_100a) only")
        self.device = device
        self._packed = None
        self._bufs = {}
        return self

    def _pack(self):
        if self._packed is not None:
            return self._packed
        dev, sd, P = self.device, self._sd, {}
        for d in range(5):
            w = sd[f"downConvs.{d}.weight"]
            P[f"dw{d}"] = w.permute(1, 2, 3, 4, 0).reshape(DOWN[d], 27, DOWN[d + 1]).contiguous().to(dev)
            P[f"db{d}"] = sd[f"downConvs.{d}.bias"].to(dev)
            w = sd[f"upConvs.{d}.weight"]
            P[f"uw{d}"] = w.permute(0, 2, 3, 4, 1).reshape(UP_IN[d], 64, UP_OUT[d]).contiguous().to(dev)
            P[f"ub{d}"] = sd[f"upConvs.{d}.bias"].to(dev)
            # split-fp16 B fragments for the mma.sync path (levels too small for it fall back to the fp32 kernels)
            P[f"uq{d}"] = ops.reg_pack_convt4(P[f"uw{d}"], UP_IN[d], UP_OUT[d]) if self.use_mma else (None, 0)
            s = sd[f"batchNorms.{d}.weight"].double() / torch.sqrt(sd[f"batchNorms.{d}.running_var"].double() + 1e-5)
            P[f"bs{d}"] = s.float().to(dev)
            P[f"bt{d}"] = (sd[f"batchNorms.{d}.bias"].double() - sd[f"batchNorms.{d}.running_mean"].double() * s
                           ).float().to(dev)
        w = torch.zeros(18, 27, 4)
        w[:, :, :3] = sd["lastConv.weight"].permute(1, 2, 3, 4, 0).reshape(18, 27, 3)
        P["lw"] = w.contiguous().to(dev)
        P["lb"] = sd["lastConv.bias"].to(dev)
        self._packed = P
        return P

    def _buffers(self, N, dims):
        key = (N, tuple(dims))
        if key not in self._bufs:
            lv = [tuple(dims)]
            for _ in range(5):
                lv.append(tuple((v + 1) // 2 for v in lv[-1]))
            cat = [torch.empty((N, UP_OUT[d] + DOWN[d]) + lv[d], dtype=torch.float32, device=self.device)
                   for d in range(5)]
            x5 = torch.empty((N, DOWN[5]) + lv[5], dtype=torch.float32, device=self.device)
            if len(self._bufs) >= 4:           # a cascade has two or three geometries; bound the cache anyway
                self._bufs.pop(next(iter(self._bufs)))
            self._bufs[key] = (lv, cat, x5)
        return self._bufs[key]

    def forward(self, A, B, out=None):
        """A, B: [N, D, H, W] float32 cuda (N = independent pairs).  Returns [N, 3, D, H, W]."""
        P = self._pack()
        N = A.shape[0]
        dims = tuple(A.shape[1:])
        lv, cat, x5 = self._buffers(N, dims)
        cat[0][:, UP_OUT[0]].copy_(A)
        cat[0][:, UP_OUT[0] + 1].copy_(B)
        for d in range(5):
            src = cat[d][:, UP_OUT[d]:]
            dst = cat[d + 1][:, UP_OUT[d + 1]:] if d < 4 else x5
            ops.reg_conv3(src, DOWN[d], P[f"dw{d}"], P[f"db{d}"], dst, DOWN[d + 1], 2, True, True)
        for d in reversed(range(5)):
            src = x5 if d == 4 else cat[d + 1]
            wpk, wexp = P[f"uq{d}"]
            ops.reg_convt4(src, UP_IN[d], P[f"uw{d}"], P[f"ub{d}"], P[f"bs{d}"], P[f"bt{d}"], cat[d][:, :UP_OUT[d]],
                           UP_OUT[d], wpk, wexp)
        if out is None:
            out = torch.empty((N, 3) + dims, dtype=torch.float32, device=self.device)
        ops.reg_conv3(cat[0], 18, P["lw"], P["lb"], out, 3, 1, False, False, 0.1)
        return out

    __call__ = forward


def tallUNet2(dimension=3, input_channels=1):
    if dimension != 3 or input_channels != 1:
        raise NotImplementedError("only tallUNet2(dimension=3, input_channels=1) is on the OAI knee path")
    return TallUNet2()
