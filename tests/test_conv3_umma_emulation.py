"""CPU tier: emulation of the index algebra of conv3s2_umma_kernel (csrc/reg_umma_down.cu) -- the space-to-depth parity
planes of reg_split_s2d_kernel, the (parity class, lattice offset) each tap reads, the zero fill outside the planes and
past an odd extent, and the clipped avg-pool residual of the epilogue -- against torch.  The hardware side (TMA boxes,
swizzle, descriptors) is covered by the GPU tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

TX, TY = 8, 16


def space_to_depth(x):
    """[cin, D, H, W] -> [8, cin, Dc, Hc, Wc]: class rz*4 + ry*2 + rx holds x[2zc + rz, 2yc + ry, 2xc + rx], zeros past
    the extent (what reg_split_s2d_kernel writes, before the hi / lo split)."""
    cin, D, H, W = x.shape
    Dc, Hc, Wc = (D + 1) // 2, (H + 1) // 2, (W + 1) // 2
    out = np.zeros((8, cin, Dc, Hc, Wc))
    for cls in range(8):
        rz, ry, rx = cls >> 2, (cls >> 1) & 1, cls & 1
        sub = x[:, rz::2, ry::2, rx::2]
        out[cls, :, :sub.shape[1], :sub.shape[2], :sub.shape[3]] = sub
    return out


def box(plane, cz, cy, cx):
    """TMA box [cin, 16, 8] of one parity plane at lattice origin (cz, cy, cx); out-of-bounds positions are zeros."""
    cin, Dc, Hc, Wc = plane.shape
    out = np.zeros((cin, TY, TX))
    if not 0 <= cz < Dc:
        return out
    for j in range(TY):
        for i in range(TX):
            y, x = cy + j, cx + i
            if 0 <= y < Hc and 0 <= x < Wc:
                out[:, j, i] = plane[:, cz, y, x]
    return out


def emulate(x_raw, w, b):
    """x_raw [cin, D, H, W], w [cout, cin, 3, 3, 3] -> Conv3d(k3, s2, p1)(leaky_relu(x)) + b + padded avg-pool residual."""
    cin, D, H, W = x_raw.shape
    cout = w.shape[0]
    xs = space_to_depth(np.where(x_raw > 0, x_raw, 0.01 * x_raw))
    Do, Ho, Wo = xs.shape[2:]
    out = np.zeros((cout, Do, Ho, Wo))
    for z in range(Do):
        for y0 in range(0, Ho, TY):
            for x0 in range(0, Wo, TX):
                acc = np.zeros((TY, TX, cout))
                for c in range(cin // 16):
                    for t in range(27):
                        kz, ky, kx = t // 9, (t // 3) % 3, t % 3
                        cls = (((kz + 1) & 1) << 2) | (((ky + 1) & 1) << 1) | ((kx + 1) & 1)
                        a = box(xs[cls, c * 16:(c + 1) * 16], z - (kz == 0), y0 - (ky == 0), x0 - (kx == 0))
                        acc += np.einsum("kji,nk->jin", a, w[:, c * 16:(c + 1) * 16, kz, ky, kx])
                for j in range(TY):
                    for i in range(TX):
                        y, xx = y0 + j, x0 + i
                        if y >= Ho or xx >= Wo:
                            continue
                        nz, ny, nx = min(2, D - 2 * z), min(2, H - 2 * y), min(2, W - 2 * xx)
                        win = x_raw[:, 2 * z:2 * z + nz, 2 * y:2 * y + ny, 2 * xx:2 * xx + nx]
                        res = np.zeros(cout)
                        res[cout - cin:] = win.reshape(cin, -1).sum(1) / (nz * ny * nx)
                        out[:, z, y, xx] = acc[j, i] + b + res
    return out


@pytest.mark.parametrize("cin,cout,dims", [(16, 32, (4, 6, 10)), (16, 32, (5, 9, 7)), (32, 64, (3, 18, 5))])
def test_down_step_schedule_equals_torch(cin, cout, dims):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(cin, *dims, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    y = F.conv3d(F.leaky_relu(x)[None], w, b, stride=2, padding=1)[0]
    pooled = F.avg_pool3d(x[None], 2, ceil_mode=True)[0]
    ref = y.clone()
    ref[cout - cin:] += pooled                      # pad_or_crop: zero channels in front of the pooled ones
    got = emulate(x.numpy(), w.numpy(), b.numpy())
    assert np.abs(got - ref.numpy()).max() < 1e-9
