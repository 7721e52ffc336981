"""CPU oracle for the ICON/GradICON registration stage -- TEST INFRASTRUCTURE ONLY.  ** parity unpinned **

The registration arithmetic of the reference lives in the un-vendored dependency `icon_registration==1.1.2`
(reference pyproject.toml:35; call sites oai_analysis/registration.py:20,25 and oai_analysis/dask_processing.py:77,85).
That package is neither under /root/reference nor installable offline, and no test in the reference asserts a
registration number (test/test_all.py:72-81,88-99 only print), so this file RESTATES the package's published
algorithm (icon_registration/networks.py::UNet2/tallUNet2, network_wrappers.py::{FunctionFromVectorField,
TwoStepRegistration, DownsampleRegistration}, mermaidlite.py::compute_warped_image_multiNC,
pretrained_models.py::OAI_knees_gradICON_model, itk_wrapper.py::{register_pair, create_itk_transform,
resampling_transform}) in plain torch, module for module, and is the checker for the CUDA path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

DOWN = [2, 16, 32, 64, 256, 512]          # tallUNet2(dimension=3): channels[0] with input_channels=1
UP_OUT = [16, 32, 64, 128, 256]           # channels[1]
UP_IN = [DOWN[d + 1] + (UP_OUT[d + 1] if d + 1 < 5 else 0) for d in range(5)]  # [48, 96, 192, 512, 512]
NET_PATHS = {  # position of the four tallUNet2 inside OAI_knees_gradICON_model().regis_net
    "phi": "netPhi.netPhi.net.netPhi.net", "psi": "netPhi.netPhi.net.netPsi.net",
    "xi": "netPhi.netPsi.net", "omega": "netPsi.net",
}
INPUT_SHAPE = (80, 192, 192)              # [BATCH, 1, 40*2, 96*2, 96*2] in pretrained_models.py


def make_unet2_state_dict(seed, last_w_std=0.02, last_b_std=0.05):
    """Random tallUNet2 weights (numpy PCG64).  icon zero-initialises lastConv (networks.UNet2.__init__), which would
    make every displacement exactly 0; SURVEY §8(d) config 2 re-randomises it so displacements are a few voxels."""
    rng = np.random.default_rng(seed)
    sd = {}

    def uni(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return torch.from_numpy(rng.uniform(-b, b, shape).astype(np.float32))

    for d in range(5):
        sd[f"downConvs.{d}.weight"] = uni((DOWN[d + 1], DOWN[d], 3, 3, 3), DOWN[d] * 27)
        sd[f"downConvs.{d}.bias"] = uni((DOWN[d + 1],), DOWN[d] * 27)
        sd[f"upConvs.{d}.weight"] = uni((UP_IN[d], UP_OUT[d], 4, 4, 4), UP_OUT[d] * 64)
        sd[f"upConvs.{d}.bias"] = uni((UP_OUT[d],), UP_OUT[d] * 64)
        sd[f"batchNorms.{d}.weight"] = torch.from_numpy((1 + 0.1 * rng.standard_normal(UP_OUT[d])).astype(np.float32))
        sd[f"batchNorms.{d}.bias"] = torch.from_numpy((0.05 * rng.standard_normal(UP_OUT[d])).astype(np.float32))
        sd[f"batchNorms.{d}.running_mean"] = torch.from_numpy(
            (0.05 * rng.standard_normal(UP_OUT[d])).astype(np.float32))
        sd[f"batchNorms.{d}.running_var"] = torch.from_numpy(
            np.exp(0.2 * rng.standard_normal(UP_OUT[d])).astype(np.float32))
        sd[f"batchNorms.{d}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["lastConv.weight"] = torch.from_numpy((last_w_std * rng.standard_normal((3, 18, 3, 3, 3))).astype(np.float32))
    sd["lastConv.bias"] = torch.from_numpy((last_b_std * rng.standard_normal(3)).astype(np.float32))
    return sd


# icon's make_network(input_shape, include_last_step=True) layout (two nested DownsampleRegistration levels):
#   TwoStep(TwoStep(Down(TwoStep(Down(phi), psi)), xi), omega)   -- quarter, half, full, full resolution
NET_PATHS_TWO_LEVEL = {
    "phi": "netPhi.netPhi.net.netPhi.net.net", "psi": "netPhi.netPhi.net.netPsi.net",
    "xi": "netPhi.netPsi.net", "omega": "netPsi.net",
}


def make_gradicon_state_dict(seed, paths=None, **kw):
    """State dict with the key layout of OAI_knees_gradICON_model().regis_net (four tallUNet2)."""
    sd = {}
    for i, (name, path) in enumerate((paths or NET_PATHS).items()):
        for k, v in make_unet2_state_dict(seed * 10 + i, **kw).items():
            sd[f"{path}.{k}"] = v
    return sd


def split_state_dict(sd):
    out = {}
    for name, path in NET_PATHS.items():
        pre = path + "."
        alt = "regis_net." + pre
        out[name] = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        if not out[name]:
            out[name] = {k[len(alt):]: v for k, v in sd.items() if k.startswith(alt)}
    return out


def pad_or_crop(x, channels):
    """icon networks.pad_or_crop: keep the first `channels` channels, or zero-pad IN FRONT."""
    y = x[:, :channels]
    if x.shape[1] < channels:
        y = F.pad(y, (0, 0, 0, 0, 0, 0, channels - x.shape[1], 0))
    return y


def unet2_forward(sd, A, B):
    """icon networks.UNet2.forward (eval mode).  A, B: [N,1,D,H,W] -> displacement [N,3,D,H,W] (already /10)."""
    x = torch.cat([A, B], 1)
    skips = []
    for d in range(5):
        skips.append(x)
        y = F.conv3d(F.leaky_relu(x), sd[f"downConvs.{d}.weight"], sd[f"downConvs.{d}.bias"], stride=2, padding=1)
        x = y + pad_or_crop(F.avg_pool3d(x, 2, ceil_mode=True), y.shape[1])
    for d in reversed(range(5)):
        y = F.conv_transpose3d(F.leaky_relu(x), sd[f"upConvs.{d}.weight"], sd[f"upConvs.{d}.bias"], stride=2, padding=1)
        x = y + F.interpolate(pad_or_crop(x, y.shape[1]), scale_factor=2, mode="trilinear", align_corners=False)
        x = F.batch_norm(x, sd[f"batchNorms.{d}.running_mean"], sd[f"batchNorms.{d}.running_var"],
                         sd[f"batchNorms.{d}.weight"], sd[f"batchNorms.{d}.bias"], training=False, eps=1e-5)
        s = skips[d]
        x = x[:, :, :s.shape[2], :s.shape[3], :s.shape[4]]
        x = torch.cat([x, s], 1)
    return F.conv3d(x, sd["lastConv.weight"], sd["lastConv.bias"], padding=1) / 10


def identity_map(shape, dtype=torch.float32):
    """mermaidlite.identity_map_multiN with spacing 1/(N-1): id[0,d,...] = index_d/(N_d-1) in [0,1]."""
    axes = [torch.arange(n, dtype=torch.float64) * (1.0 / (n - 1)) for n in shape]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), 0)[None].to(dtype)


def sample(I, coords):
    """RegistrationModule.as_function(I)(coords) = compute_warped_image_multiNC(I, coords, spacing, 1):
    grid_sample(bilinear, border, align_corners=True) at 2c-1 with channels reversed to x,y,z."""
    grid = (coords * 2 - 1)[:, [2, 1, 0]].permute(0, 2, 3, 4, 1)
    return F.grid_sample(I, grid, mode="bilinear", padding_mode="border", align_corners=True)


def regis_net_forward(nets, A, B):
    """TwoStep(TwoStep(Downsample(TwoStep(phi, psi)), xi), omega).forward(A, B): returns the four displacement
    tensors in application order for B.4 and nothing else (closures are re-evaluated exactly as the reference does)."""
    id_full = identity_map(A.shape[2:], A.dtype)
    A_lo, B_lo = F.avg_pool3d(A, 2, ceil_mode=True), F.avg_pool3d(B, 2, ceil_mode=True)
    id_lo = identity_map(A_lo.shape[2:], A.dtype)
    u_phi = unet2_forward(nets["phi"], A_lo, B_lo)
    t_phi = lambda c: c + sample(u_phi, c)                                   # noqa: E731
    u_psi = unet2_forward(nets["psi"], sample(A_lo, t_phi(id_lo)), B_lo)
    t_psi = lambda c: c + sample(u_psi, c)                                   # noqa: E731
    low = lambda c: t_phi(t_psi(c))                                          # noqa: E731
    u_xi = unet2_forward(nets["xi"], sample(A, low(id_full)), B)
    t_xi = lambda c: c + sample(u_xi, c)                                     # noqa: E731
    hires = lambda c: low(t_xi(c))                                           # noqa: E731
    u_omega = unet2_forward(nets["omega"], sample(A, hires(id_full)), B)
    return dict(phi=u_phi, psi=u_psi, xi=u_xi, omega=u_omega)


# ---- generic module tree (network_wrappers.py restated as closures, module for module) -------------------------
_UNET_KEYS = ("downConvs.", "upConvs.", "batchNorms.", "lastConv.")


def tree_from_state_dict(sd):
    """(tree, {path: unet state dict}) from a regis_net state dict: 'netPhi'/'netPsi' are TwoStepRegistration children,
    a trailing 'net' is FunctionFromVectorField.net, any other 'net' is DownsampleRegistration.net."""
    groups = {}
    for k, v in sd.items():
        k = k[len("regis_net."):] if k.startswith("regis_net.") else k
        cut = min(k.find(h) for h in _UNET_KEYS if h in k)
        groups.setdefault(k[:cut - 1], {})[k[cut:]] = v

    def build(prefix, rel):
        if rel == [("net",)]:
            return ("ffvf", ".".join(prefix + ("net",)))
        if {r[0] for r in rel} == {"net"}:
            return ("down", build(prefix + ("net",), [r[1:] for r in rel]))
        return ("twostep", build(prefix + ("netPhi",), [r[1:] for r in rel if r[0] == "netPhi"]),
                build(prefix + ("netPsi",), [r[1:] for r in rel if r[0] == "netPsi"]))
    return build((), [tuple(p.split(".")) for p in sorted(groups)]), groups


def tree_forward(node, nets, A, B):
    """RegistrationModule.forward(A, B) of a tree node: returns the transform closure t(coords, is_identity_map)."""
    if node[0] == "ffvf":            # FunctionFromVectorField
        u = unet2_forward(nets[node[1]], A, B)
        return lambda c, ident=False: c + u if (ident and c.shape == u.shape) else c + sample(u, c)
    if node[0] == "down":            # DownsampleRegistration
        return tree_forward(node[1], nets, F.avg_pool3d(A, 2, ceil_mode=True), F.avg_pool3d(B, 2, ceil_mode=True))
    phi = tree_forward(node[1], nets, A, B)          # TwoStepRegistration
    warped = sample(A, phi(identity_map(A.shape[2:], A.dtype), True))
    psi = tree_forward(node[2], nets, warped, B)
    return lambda c, ident=False: phi(psi(c, ident))


def register_pair_maps_tree(sd, image_A, image_B, shape=INPUT_SHAPE):
    """register_pair_maps for any module tree, driven by the state dict's key paths."""
    tree, nets = tree_from_state_dict(sd)
    A, B = resize_to_network(np.asarray(image_A), shape), resize_to_network(np.asarray(image_B), shape)
    ident = identity_map(shape)
    with torch.no_grad():
        return tree_forward(tree, nets, A, B)(ident, True), tree_forward(tree, nets, B, A)(ident, True)


def final_map(u, shape, dtype=torch.float32):
    """model.phi_AB(model.identity_map): identity shortcut on omega, then xi, psi, phi."""
    c = identity_map(shape, dtype) + u["omega"]
    c = c + sample(u["xi"], c)
    c = c + sample(u["psi"], c)
    return c + sample(u["phi"], c)


def resize_to_network(img, shape=INPUT_SHAPE):
    """itk_wrapper.register_pair: F.interpolate(size=identity_map.shape[2:], trilinear, align_corners=False)."""
    t = torch.as_tensor(np.asarray(img), dtype=torch.float32)[None, None]
    return F.interpolate(t, size=tuple(shape), mode="trilinear", align_corners=False)


def register_pair_maps(sd, image_A, image_B, shape=INPUT_SHAPE):
    """register_pair up to (phi_AB, phi_BA) as [1,3,*shape] coordinate maps in [0,1]."""
    a, b = np.asarray(image_A), np.asarray(image_B)
    assert a.max() != a.min() and b.max() != b.min()
    nets = split_state_dict(sd)
    A, B = resize_to_network(a, shape), resize_to_network(b, shape)
    with torch.no_grad():
        phi_AB = final_map(regis_net_forward(nets, A, B), shape)
        phi_BA = final_map(regis_net_forward(nets, B, A), shape)
    return phi_AB, phi_BA


def displacement_field_xyz(phi, shape=INPUT_SHAPE):
    """create_itk_transform: disp = (phi - id) * (N - 1), components reversed to x,y,z, layout [D,H,W,3], float64."""
    disp = (phi - identity_map(shape, phi.dtype))[0]
    scale = torch.tensor([n - 1 for n in shape], dtype=phi.dtype).view(3, 1, 1, 1)
    disp = (disp * scale).double().numpy()
    return np.ascontiguousarray(disp[::-1].transpose(1, 2, 3, 0))
