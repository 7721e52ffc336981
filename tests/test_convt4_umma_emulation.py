"""CPU tier: emulation of the index algebra of convt4_umma_kernel (csrc/reg_umma.cu) -- the parity-class / shift
formulation of ConvTranspose3d(k4, s2, p1), the slot order of the stacked weight rows, the column span an input slice
updates, the overwrite-on-first-touch rule of a unit's first pass and the epilogue's column -> (class, channel) map --
against torch.  The hardware side (TMA boxes, swizzle, descriptors) is covered by the GPU tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

TX, TY = 8, 16


def pack_rows(w, Cn, h, c, j):
    """Rows of weight block (h, c, j) as the pack kernel orders them: [16 * Cn, 16] (before the hi / lo split)."""
    cin, cout = w.shape[0], w.shape[1]
    rows = np.zeros((16 * Cn, 16), dtype=np.float64)
    sy, sx = j // 3 - 1, j % 3 - 1
    for n in range(16 * Cn):
        slot, co = n // Cn, h * Cn + n % Cn
        if slot < 4:
            sd, cls = 1, 4 + slot
        elif slot < 12:
            sd, cls = 0, slot - 4
        else:
            sd, cls = -1, slot - 12
        kd, kh, kw = (cls >> 2) + 1 - 2 * sd, ((cls >> 1) & 1) + 1 - 2 * sy, (cls & 1) + 1 - 2 * sx
        if 0 <= kd < 4 and 0 <= kh < 4 and 0 <= kw < 4:
            rows[n] = w[c * 16:(c + 1) * 16, co, kd, kh, kw]
    return rows


def emulate(x, w, out_dims):
    """x [cin, D, H, W] (already leaky-relu'd), w [cin, cout, 4, 4, 4] -> [cout, *out_dims] by the kernel's schedule."""
    cin, Di, Hi, Wi = x.shape
    cout = w.shape[1]
    Cn = 32 if cout >= 32 else 16
    R = 256 // (8 * Cn)
    out = np.zeros((cout,) + tuple(out_dims))
    xp = np.zeros((cin, Di, Hi + TY + 2, Wi + TX + 2))   # zero fill outside the volume = TMA out-of-bounds fill
    xp[:, :, 1:Hi + 1, 1:Wi + 1] = x
    for h in range(cout // Cn):
        for z0 in range(0, Di, R):
            rd = min(R, Di - z0)
            for y0 in range(0, Hi, TY):
                for x0 in range(0, Wi, TX):
                    tmem = np.full((128, 256), np.nan)   # stale contents of the TMEM half
                    ncols, hw = rd * 8 * Cn, 0
                    zlo, zhi = max(0, z0 - 1), min(Di - 1, z0 + rd)
                    for c in range(cin // 16):
                        for j in range(9):
                            B = pack_rows(w, Cn, h, c, j)
                            for zi in range(zlo, zhi + 1):
                                start = (zi - 1 - z0) * 8 * Cn + 4 * Cn
                                c0, c1 = max(start, 0), min(start + 16 * Cn, ncols)
                                if c1 <= c0:
                                    continue
                                # A: box origin (x0 - 1, y0 - 1), shifted (j // 3, j % 3) rows / columns
                                box = xp[c * 16:(c + 1) * 16, zi, y0 + j // 3:y0 + j // 3 + TY, x0 + j % 3:x0 + j % 3 + TX]
                                A = box.reshape(16, 128).T                      # [m = ty * 8 + tx, k]
                                prod = A @ B[c0 - start:c1 - start].T           # [128, c1 - c0]
                                if c == 0 and j == 0:
                                    mid = min(max(hw, c0), c1)
                                    tmem[:, c0:mid] += prod[:, :mid - c0]
                                    tmem[:, mid:c1] = prod[:, mid - c0:]
                                    hw = max(hw, c1)
                                else:
                                    tmem[:, c0:c1] += prod
                    assert not np.isnan(tmem[:, :ncols]).any()
                    for m in range(128):
                        y, xx = y0 + m // 8, x0 + m % 8
                        if y >= Hi or xx >= Wi:
                            continue
                        for a in range(rd):
                            for cls in range(8):
                                zo, yo, xo = 2 * (z0 + a) + (cls >> 2), 2 * y + ((cls >> 1) & 1), 2 * xx + (cls & 1)
                                if zo < out_dims[0] and yo < out_dims[1] and xo < out_dims[2]:
                                    col = a * 8 * Cn + cls * Cn
                                    out[h * Cn:(h + 1) * Cn, zo, yo, xo] = tmem[m, col:col + Cn]
    return out


@pytest.mark.parametrize("cin,cout,dims,crop", [
    (16, 16, (3, 5, 9), (6, 10, 17)),
    (32, 32, (2, 17, 4), (3, 34, 8)),
    (16, 64, (1, 3, 3), (2, 6, 6)),
])
def test_umma_schedule_equals_conv_transpose(cin, cout, dims, crop):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(cin, *dims, generator=g, dtype=torch.float64)
    w = torch.randn(cin, cout, 4, 4, 4, generator=g, dtype=torch.float64)
    ref = F.conv_transpose3d(x[None], w, None, stride=2, padding=1)[0][:, :crop[0], :crop[1], :crop[2]]
    got = emulate(x.numpy(), w.numpy(), crop)
    assert np.abs(got - ref.numpy()).max() < 1e-9
