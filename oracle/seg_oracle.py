"""CPU oracle for the segmentation stage -- TEST INFRASTRUCTURE ONLY.

A plain torch/numpy restatement of the reference's patch-wise 3-D UNet prediction path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the product
path (oai_analysis_2_b200/) never does.

Pinned: tests/golden/seg_*.npz hold outputs of the reference's OWN code
(oai_analysis/segmentation/segmenter.py::Segmenter3DInPatchClassWise.segment, run unmodified through a stub `itk`
by tests/golden/make_golden.py in the build container); tests/test_oracle_seg.py checks this restatement against
them.  The reference's only quantitative test (test/test_all.py:32-33, sum|dprob| < 12 vs golden NIfTIs with the
pretrained weights) needs release tarballs that are not available offline.
"""
import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------------------
# Partition (reference: oai_analysis/segmentation/image_transforms.py:388-455)
# ---------------------------------------------------------------------------------------------------------------


def tiling_geometry(image_shape_zyx, patch_size_xyz, overlap_xyz):
    """image_transforms.py:389-391 (x,y,z -> z,y,x flip) and :404-406 (effective size, grid, padding)."""
    tile = np.asarray(patch_size_xyz)[::-1].astype(int)
    overlap = np.asarray(overlap_xyz)[::-1].astype(int)
    image = np.asarray(image_shape_zyx).astype(int)
    effective = tile - 2 * overlap
    grid = np.ceil(image / effective).astype(int)
    padded = effective * grid + 2 * overlap - image
    return dict(tile=tile, overlap=overlap, image=image, effective=effective, grid=grid, padded=padded)


def partition(volume, patch_size_xyz, overlap_xyz):
    """image_transforms.py:408-446: reflect pad, then slice tiles in i(z)-major order.  Returns [T,1,d,h,w] float32."""
    g = tiling_geometry(volume.shape, patch_size_xyz, overlap_xyz)
    ov, pad, eff, tile = g["overlap"], g["padded"], g["effective"], g["tile"]
    padded = np.pad(volume, [(ov[a], pad[a] - ov[a]) for a in range(3)], mode="reflect")
    tiles = []
    for i in range(g["grid"][0]):
        for j in range(g["grid"][1]):
            for k in range(g["grid"][2]):
                tiles.append(padded[i * eff[0]:i * eff[0] + tile[0], j * eff[1]:j * eff[1] + tile[1],
                                    k * eff[2]:k * eff[2] + tile[2]])
    return torch.from_numpy(np.stack(tiles, 0)[:, None].astype(np.float32)), g


def assemble(tiles, g, crop_size_xyz):
    """image_transforms.py:492-513 (non-vote branch): place tile interiors, trim, zero the border shell; float64."""
    tiles = np.asarray(tiles)
    ov, eff, tile, grid, image = g["overlap"], g["effective"], g["tile"], g["grid"], g["image"]
    out = np.zeros(eff * grid)
    for i in range(grid[0]):
        for j in range(grid[1]):
            for k in range(grid[2]):
                t = tiles[(i * grid[1] + j) * grid[2] + k]
                out[i * eff[0]:(i + 1) * eff[0], j * eff[1]:(j + 1) * eff[1], k * eff[2]:(k + 1) * eff[2]] = \
                    t[ov[0]:tile[0] - ov[0], ov[1]:tile[1] - ov[1], ov[2]:tile[2] - ov[2]]
    out = out[:image[0], :image[1], :image[2]]
    if crop_size_xyz:
        cx, cy, cz = crop_size_xyz  # reference indexes crop_size[2], [0], [1] for z, y(!), x(!): :511
        c = np.zeros(out.shape)
        c[cz:-cz, cx:-cx, cy:-cy] = out[cz:-cz, cx:-cx, cy:-cy]
        out = c
    return out


# ---------------------------------------------------------------------------------------------------------------
# UNet (reference: oai_analysis/segmentation/networks.py:38-149)
# ---------------------------------------------------------------------------------------------------------------

# name, kind ("c" Conv3d k3 p1 | "t3" ConvTranspose3d k3 s1 p1 | "t2" ConvTranspose3d k2 s2), cin, cout
UNET_LAYERS = [
    ("ec0", "c", None, 32), ("ec1", "c", 32, 64), ("ec2", "c", 64, 64), ("ec3", "c", 64, 128),
    ("ec4", "c", 128, 128), ("ec5", "c", 128, 256), ("ec6", "c", 256, 256), ("ec7", "c", 256, 512),
    ("dc9", "t2", 512, 512), ("dc8", "t3", 768, 256), ("dc7", "t3", 256, 256), ("dc6", "t2", 256, 256),
    ("dc5", "t3", 384, 128), ("dc4", "t3", 128, 128), ("dc3", "t2", 128, 128), ("dc2", "t3", 192, 64),
    ("dc1", "t3", 64, 64),
]


def unet_layer_table(in_channels):
    return [(n, k, in_channels if ci is None else ci, co) for n, k, ci, co in UNET_LAYERS]


def make_unet_state_dict(seed, in_channels=1, n_classes=2, bias=True, BN=True, trained_like=True, head_gain=1.0,
                         head_bias=None):
    """Deterministic (numpy PCG64) weights with the reference's state_dict keys and shapes.

    Conv weights follow the reference's xavier_normal_ statistics (networks.py:71-78, std = sqrt(2/(fan_in+fan_out)));
    with trained_like=True biases and BatchNorm affine/running statistics are non-trivial so BN folding is exercised.
    Plain random init clusters the sigmoid output at 0.497 +- 0.002, which makes a 0.5-threshold
    mask meaningless; head_gain / head_bias (constants calibrated once per test case, see calibrate_head) rescale the
    zero-mean dc0 filter so interior logits are ~N(0, 2^2) and the mask is a non-trivial structure.
    """
    rng = np.random.default_rng(seed)
    sd = {}

    def conv_w(shape):
        k3 = int(np.prod(shape[2:]))
        std = (2.0 / ((shape[0] + shape[1]) * k3)) ** 0.5
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))

    for name, kind, ci, co in unet_layer_table(in_channels):
        k = 2 if kind == "t2" else 3
        shape = (co, ci, k, k, k) if kind == "c" else (ci, co, k, k, k)
        sd[f"{name}.0.weight"] = conv_w(shape)
        if bias:
            b = rng.standard_normal(co) * 0.05 if trained_like else np.zeros(co)
            sd[f"{name}.0.bias"] = torch.from_numpy(b.astype(np.float32))
        if BN:
            if trained_like:
                gamma, beta = 1 + 0.1 * rng.standard_normal(co), 0.05 * rng.standard_normal(co)
                mean, var = 0.05 * rng.standard_normal(co), np.exp(0.2 * rng.standard_normal(co))
            else:
                gamma, beta, mean, var = np.ones(co), np.zeros(co), np.zeros(co), np.ones(co)
            sd[f"{name}.1.weight"] = torch.from_numpy(gamma.astype(np.float32))
            sd[f"{name}.1.bias"] = torch.from_numpy(beta.astype(np.float32))
            sd[f"{name}.1.running_mean"] = torch.from_numpy(mean.astype(np.float32))
            sd[f"{name}.1.running_var"] = torch.from_numpy(var.astype(np.float32))
            sd[f"{name}.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    w0 = conv_w((n_classes, 64, 1, 1, 1))
    if head_bias is not None:
        w0 = w0 - w0.mean(dim=1, keepdim=True)
    sd["dc0.weight"] = w0 * head_gain
    if bias:
        if head_bias is not None:
            b = np.asarray(head_bias, dtype=np.float64)
        else:
            b = rng.standard_normal(n_classes) * 0.05 if trained_like else np.zeros(n_classes)
        sd["dc0.bias"] = torch.from_numpy(b.astype(np.float32))
    return sd


def calibrate_head(seed, volume, patch_size_xyz, overlap_xyz, BN=True, bias=True, n_tiles=4, target_std=2.0):
    """One-off helper that produced the head_gain/head_bias constants recorded in tests/golden/make_golden.py and
    bench.py: gain and per-class bias that make the interior logits of the first n_tiles tiles ~N(0, target_std^2)."""
    sd = make_unet_state_dict(seed, 1, 2, bias, BN, True, 1.0, head_bias=[0.0, 0.0])
    tiles, g = partition(np.asarray(volume), patch_size_xyz, overlap_xyz)
    ov = g["overlap"]
    idx = np.linspace(0, tiles.shape[0] - 1, n_tiles).astype(int)
    with torch.no_grad():
        lo = unet_forward(sd, tiles[idx], BN)[:, :, ov[0]:-ov[0], ov[1]:-ov[1], ov[2]:-ov[2]]
    std = float(lo.std())
    gain = target_std / std
    med = [float(lo[:, c].median()) for c in range(lo.shape[1])]
    return gain, [-gain * m for m in med]


def _block(sd, name, kind, x, BN):
    w, b = sd[f"{name}.0.weight"], sd.get(f"{name}.0.bias")
    if kind == "c":
        x = F.conv3d(x, w, b, padding=1)
    elif kind == "t3":
        x = F.conv_transpose3d(x, w, b, stride=1, padding=1)
    else:
        x = F.conv_transpose3d(x, w, b, stride=2)
    if BN:
        x = F.batch_norm(x, sd[f"{name}.1.running_mean"], sd[f"{name}.1.running_var"], sd[f"{name}.1.weight"],
                         sd[f"{name}.1.bias"], training=False, eps=1e-5)
    return F.relu(x)


def unet_forward(sd, x, BN=True):
    """networks.py:109-149 (eval mode).  x: [N, Cin, D, H, W] float tensor -> logits [N, n_classes, D, H, W]."""
    kinds = {n: k for n, k, _, _ in UNET_LAYERS}
    blk = lambda n, t: _block(sd, n, kinds[n], t, BN)  # noqa: E731
    syn0 = blk("ec1", blk("ec0", x))
    syn1 = blk("ec3", blk("ec2", F.max_pool3d(syn0, 2)))
    syn2 = blk("ec5", blk("ec4", F.max_pool3d(syn1, 2)))
    e7 = blk("ec7", blk("ec6", F.max_pool3d(syn2, 2)))
    d7 = blk("dc7", blk("dc8", torch.cat((blk("dc9", e7), syn2), 1)))
    d4 = blk("dc4", blk("dc5", torch.cat((blk("dc6", d7), syn1), 1)))
    d1 = blk("dc1", blk("dc2", torch.cat((blk("dc3", d4), syn0), 1)))
    return F.conv3d(d1, sd["dc0.weight"], sd.get("dc0.bias"))


def segment(volume, sd, patch_size_xyz=(128, 128, 32), overlap_xyz=(16, 16, 8), batch_size=4, BN=True,
            output_prob=True, dtype=torch.float32, return_tiles=False):
    """segmenter.py:100-131: tile -> batched forward -> sigmoid (-> >0.5) -> assemble FC (ch 0) and TC (ch 1)."""
    tiles, g = partition(np.asarray(volume), patch_size_xyz, overlap_xyz)
    sdd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
    outs = []
    with torch.no_grad():
        for i in range(0, tiles.shape[0], batch_size):
            outs.append(unet_forward(sdd, tiles[i:i + batch_size].to(dtype), BN).float())
        pred = torch.sigmoid(torch.cat(outs, 0))
        if not output_prob:
            pred = pred > 0.5
    fc = assemble(pred[:, 0].numpy(), g, overlap_xyz)
    tc = assemble(pred[:, 1].numpy(), g, overlap_xyz)
    if return_tiles:
        return fc, tc, pred
    return fc, tc


def synthetic_knee(shape_zyx, seed, n_blobs=64):
    """SURVEY §8(d) config 1 input: sum of random Gaussian blobs + 0.05*U noise, min-max scaled to [0,1] float32."""
    rng = np.random.default_rng(seed)
    D, H, W = shape_zyx
    z, y, x = np.meshgrid(np.arange(D, dtype=np.float32), np.arange(H, dtype=np.float32),
                          np.arange(W, dtype=np.float32), indexing="ij", sparse=True)
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
