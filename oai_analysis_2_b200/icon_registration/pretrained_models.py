"""icon_registration.pretrained_models.OAI_knees_gradICON_model on B200.

The reference builds GradientICON(TwoStep(TwoStep(Downsample(TwoStep(phi, psi)), xi), omega), similarity, lmbda) and
at inference only uses the four tallUNet2 and the closures composing them; the loss terms GradientICON.forward also
computes are discarded by register_pair, so they are not evaluated here."""
import os

import numpy as np
import torch

from .. import ops
from .networks import TallUNet2

NET_PATHS = {"phi": "netPhi.netPhi.net.netPhi.net", "psi": "netPhi.netPhi.net.netPsi.net",
             "xi": "netPhi.netPsi.net", "omega": "netPsi.net"}
INPUT_SHAPE = [1, 1, 80, 192, 192]
WEIGHTS_ENV = "OAI_B200_GRADICON_WEIGHTS"


class GradICONModel:
    """Inference-side equivalent of the object OAI_knees_gradICON_model() returns: exposes .identity_map,
    .assign_identity_map, .to/.cuda/.eval, .regis_net state loading, __call__(A, B), .phi_AB / .phi_BA."""

    def __init__(self):
        self.nets = {k: TallUNet2() for k in NET_PATHS}
        self.device = torch.device("cpu")
        self.input_shape = list(INPUT_SHAPE)
        self._identity = None
        self.phi_AB_vectorfield = None
        self.phi_BA_vectorfield = None

    # -- module-like surface
    def assign_identity_map(self, input_shape):
        self.input_shape = [1, 1] + [int(v) for v in input_shape[2:]]
        self._identity = None

    @property
    def identity_map(self):
        if self._identity is None:
            D, H, W = self.input_shape[2:]
            axes = [torch.arange(n, dtype=torch.float64) * (1.0 / (n - 1)) for n in (D, H, W)]
            self._identity = torch.stack(torch.meshgrid(*axes, indexing="ij"), 0)[None].float().to(self.device)
        return self._identity

    def to(self, device):
        self.device = torch.device(device)
        for n in self.nets.values():
            n.to(self.device)
        self._identity = None
        return self

    def cuda(self):
        return self.to("cuda")

    def eval(self):
        return self

    def state_dict(self):
        sd = {}
        for name, path in NET_PATHS.items():
            sd.update(self.nets[name].state_dict(prefix=f"regis_net.{path}."))
        return sd

    def load_state_dict(self, sd, strict=False):
        """Accepts the reference checkpoint layout (keys of regis_net, with or without the 'regis_net.' prefix)."""
        for name, path in NET_PATHS.items():
            for pre in (path + ".", "regis_net." + path + "."):
                sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
                if sub:
                    self.nets[name].load_state_dict(sub, strict=True)
                    break
            else:
                if strict:
                    raise RuntimeError(f"GradICON state dict has no weights for {path}")

    # -- inference
    def forward(self, image_A, image_B):
        """image_A/B: [1,1,D,H,W] (or [D,H,W]) float32 cuda at the network resolution.  Runs both directions
        (batched through each UNet) and stores phi_AB / phi_BA evaluated on the identity map."""
        A = image_A.reshape(image_A.shape[-3:]).contiguous().float()
        B = image_B.reshape(image_B.shape[-3:]).contiguous().float()
        full = tuple(A.shape)
        if list(full) != self.input_shape[2:]:
            raise ValueError(f"images must be resized to the network shape {self.input_shape[2:]}, got {list(full)}")
        src = torch.stack((A, B))                    # direction 0 registers A->B, direction 1 B->A
        tgt = torch.stack((B, A))
        lo_src = ops.avgpool2_ceil(src)
        lo_tgt = torch.stack((lo_src[1], lo_src[0]))
        lo = tuple(lo_src.shape[1:])
        n = self.nets
        u_phi = n["phi"](lo_src, lo_tgt)
        warped = torch.empty_like(lo_src)
        for k in range(2):
            ops.compose(lo, [u_phi[k]], False, lo_src[k], want_phi=False, img_out=warped[k])
        u_psi = n["psi"](warped, lo_tgt)
        warped = torch.empty_like(src)
        for k in range(2):
            ops.compose(full, [u_psi[k], u_phi[k]], False, src[k], want_phi=False, img_out=warped[k])
        u_xi = n["xi"](warped, tgt)
        for k in range(2):
            ops.compose(full, [u_xi[k], u_psi[k], u_phi[k]], False, src[k], want_phi=False, img_out=warped[k])
        u_omega = n["omega"](warped, tgt)
        maps = [ops.compose(full, [u_omega[k], u_xi[k], u_psi[k], u_phi[k]], True)[0] for k in range(2)]
        self.displacements = dict(phi=u_phi, psi=u_psi, xi=u_xi, omega=u_omega)
        self.phi_AB_vectorfield, self.phi_BA_vectorfield = maps[0][None], maps[1][None]
        return self.phi_AB_vectorfield, self.phi_BA_vectorfield

    __call__ = forward

    def _eval_on_identity(self, field, coords):
        if field is None:
            raise RuntimeError("call the model on an image pair first")
        if coords is not self.identity_map and tuple(coords.shape) != tuple(self.identity_map.shape):
            raise NotImplementedError("phi_AB / phi_BA are evaluated on the model's identity map (register_pair usage)")
        return field

    def phi_AB(self, coords):
        return self._eval_on_identity(self.phi_AB_vectorfield, coords)

    def phi_BA(self, coords):
        return self._eval_on_identity(self.phi_BA_vectorfield, coords)


def OAI_knees_gradICON_model(pretrained=True, weights_path=None):
    """pretrained_models.OAI_knees_gradICON_model.  The reference downloads the checkpoint from a GitHub release;
    here `weights_path` (or $OAI_B200_GRADICON_WEIGHTS) must point at that file when pretrained=True."""
    net = GradICONModel()
    net.assign_identity_map(INPUT_SHAPE)
    if pretrained:
        path = weights_path or os.environ.get(WEIGHTS_ENV)
        if not path or not os.path.isfile(path):
            raise FileNotFoundError(
                "pretrained GradICON knee weights not found: pass weights_path= or set $%s (no network access to the "
                "release the reference downloads from)" % WEIGHTS_ENV)
        net.load_state_dict(torch.load(path, map_location="cpu", weights_only=False), strict=False)
    if torch.cuda.is_available():
        net.to("cuda")
    net.eval()
    return net
