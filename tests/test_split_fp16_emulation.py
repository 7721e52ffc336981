"""CPU tier: numpy emulation of the split-fp16 arithmetic used by convt4_mma_kernel / stem_mma_kernel
(x = hi + lo in fp16, hi*hi + lo*hi + hi*lo accumulated in fp32) and of the transposed-conv parity/tap mapping the
kernel iterates over.  Documents the accuracy claim (fp32-level, ~2^-22 relative) without a GPU."""
import numpy as np
import torch
import torch.nn.functional as F


def split(x, scale=1.0):
    x = np.asarray(x, dtype=np.float32) * np.float32(scale)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def test_three_term_split_product_is_fp32_level():
    rng = np.random.default_rng(0)
    K = 384
    a = rng.standard_normal((256, K)).astype(np.float32) * 2.0
    w = (rng.standard_normal((K, 16)) * 0.05).astype(np.float32)
    wexp = int(13 - np.floor(np.log2(np.abs(w).max())))           # ops.reg_pack_convt4's scale rule
    ah, al = split(a)
    wh, wl = split(w, 2.0 ** wexp)
    ws = w * np.float32(2.0 ** wexp)
    assert np.abs(ws - wh - wl).max() <= 2.0 ** -23 * np.abs(ws).max()   # hi + lo carries ~22 bits of the big weights
    acc = (ah @ wh + al @ wh + ah @ wl).astype(np.float32) * np.float32(2.0 ** -wexp)
    ref = a.astype(np.float64) @ w.astype(np.float64)
    fp32 = a @ w
    scale = np.abs(ref).max()
    e_split, e_fp32, e_fp16 = (np.abs(acc - ref).max() / scale, np.abs(fp32 - ref).max() / scale,
                               np.abs(ah @ split(w)[0] - ref).max() / scale)
    assert e_split < 2e-6 and e_split < 40 * max(e_fp32, 1e-8)     # a few fp32 ulps
    assert e_fp16 > 50 * e_split                                   # plain fp16 operands are ~2^-11


def test_parity_class_tap_mapping_reproduces_conv_transpose():
    """o = 2 i - 1 + k: class p = 0 uses (k=1, i=q), (k=3, i=q-1); p = 1 uses (k=2, i=q), (k=0, i=q+1) -- the
    (class, tap) <-> neighbour-shift table of convt4_mma_kernel (sh_par / sh_tap), checked against torch in 1-D x 3."""
    def sh_ncomb(s):
        return 2 if s == 1 else 1

    def sh_par(s, i):
        return 0 if s == 0 else (1 if s == 2 else i)

    def sh_tap(s, i):
        return 3 if s == 0 else (0 if s == 2 else (1 if i == 0 else 2))

    g = torch.Generator().manual_seed(1)
    cin, cout, dims = 3, 2, (3, 4, 5)
    x = torch.randn(1, cin, *dims, generator=g, dtype=torch.float64)
    w = torch.randn(cin, cout, 4, 4, 4, generator=g, dtype=torch.float64)
    ref = F.conv_transpose3d(x, w, stride=2, padding=1)[0]
    xp = F.pad(x[0], (1, 1, 1, 1, 1, 1))                           # zero halo = the conv's implicit padding
    out = torch.zeros_like(ref)
    visits = 0
    for sz in range(3):
        for sy in range(3):
            for sx in range(3):
                for iz in range(sh_ncomb(sz)):
                    for iy in range(sh_ncomb(sy)):
                        for ix in range(sh_ncomb(sx)):
                            pz, py, px = sh_par(sz, iz), sh_par(sy, iy), sh_par(sx, ix)
                            kz, ky, kx = sh_tap(sz, iz), sh_tap(sy, iy), sh_tap(sx, ix)
                            a = xp[:, sz:sz + dims[0], sy:sy + dims[1], sx:sx + dims[2]]     # input at q + s - 1
                            out[:, pz::2, py::2, px::2] += torch.einsum("izyx,io->ozyx", a, w[:, :, kz, ky, kx])
                            visits += 1
    assert visits == 64
    assert (out - ref).abs().max() < 1e-12


def test_split_k_gemm_view_of_the_strided_conv():
    """deep_gemm_kernel<0> treats Conv3d k3 s2 p1 as one GEMM per tap: A[m][ci] = x[ci][2 o - 1 + k] (zero outside),
    partials summed over (tap, channel range) in a fixed order.  Emulated in numpy against torch."""
    g = torch.Generator().manual_seed(2)
    cin, cout, dims = 6, 5, (5, 6, 7)
    x = torch.randn(1, cin, *dims, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, stride=2, padding=1)[0]
    Do, Ho, Wo = ref.shape[1:]
    xp = F.pad(x[0], (1, 1, 1, 1, 1, 1))
    nsplit, per = 2, 3                                            # two channel ranges per tap, like ci_per_split
    partial = torch.zeros(27 * nsplit, cout, Do, Ho, Wo, dtype=torch.float64)
    for tap in range(27):
        kd, kh, kw = tap // 9, (tap // 3) % 3, tap % 3
        a = xp[:, kd:kd + 2 * Do:2, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2]   # input at 2 o - 1 + k
        for s in range(nsplit):
            cs = slice(s * per, (s + 1) * per)
            partial[tap * nsplit + s] = torch.einsum("izyx,oi->ozyx", a[cs], w[:, cs, kd, kh, kw])
    out = torch.zeros_like(ref)
    for k in range(27 * nsplit):                                  # the reduce kernel's fixed order
        out += partial[k]
    assert (out - ref).abs().max() < 1e-12
