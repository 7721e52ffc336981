"""CPU tier: emulation of mma.sync.m16n8k16 fragment layouts (PTX ISA: A row-major 16x16, B col-major 16x8, C 16x8)
and of the index algebra stem_mma_kernel builds on them: taps as the K axis (27 padded to 32), the per-thread tap
table, and the permuted n index (column n of n-tile nt = channel 8 (n / 2) + 2 nt + n % 2) that makes a lane's C
fragments 8 consecutive channels."""
import numpy as np


def mma_16816(a_frag, b_frag, c_frag):
    """a_frag [32][4][2], b_frag [32][2][2], c_frag [32][4] -> c_frag + A @ B with the PTX fragment layouts."""
    A, B = np.zeros((16, 16)), np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * t + e] = a_frag[lane][0][e]
            A[g + 8, 2 * t + e] = a_frag[lane][1][e]
            A[g, 2 * t + 8 + e] = a_frag[lane][2][e]
            A[g + 8, 2 * t + 8 + e] = a_frag[lane][3][e]
            B[2 * t + e, g] = b_frag[lane][0][e]
            B[2 * t + 8 + e, g] = b_frag[lane][1][e]
    C = A @ B
    out = np.array(c_frag, dtype=np.float64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane] += [C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1]]
    return out


def test_stem_fragment_algebra_reproduces_the_27_tap_conv():
    rng = np.random.default_rng(0)
    SW, SH, C0 = 130, 6, 32
    win = rng.standard_normal((3, SH, SW))          # [slice d-1, d, d+1][row][x], halo included
    w = rng.standard_normal((27, C0))               # [tap][channel], tap = (kd*3 + kh)*3 + kw
    bias = rng.standard_normal(C0)
    ly, j = 2, 3                                    # warp's row, m-tile within the row
    acc = np.zeros((4, 32, 4))                      # [nt][lane][4]
    for lane in range(32):
        t = lane & 3
        for nt in range(4):
            acc[nt][lane] = [bias[8 * t + 2 * nt], bias[8 * t + 2 * nt + 1]] * 2
    for s in range(2):                              # two k-steps of 16 taps
        a_frag = np.zeros((32, 4, 2))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for h in range(2):                      # column half
                for e in range(2):
                    k = 16 * s + 8 * h + 2 * t + e
                    kk = k if k < 27 else 0          # padded taps read a valid sample; their weights are zero
                    kd, rel = kk // 9, ((kk // 3) % 3) * SW + kk % 3
                    base = win[kd].reshape(-1)
                    a_frag[lane][2 * h][e] = base[rel + ly * SW + g + 16 * j]          # row g
                    a_frag[lane][2 * h + 1][e] = base[rel + ly * SW + g + 8 + 16 * j]  # row g + 8
        for nt in range(4):
            b_frag = np.zeros((32, 2, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                co = 8 * (g >> 1) + 2 * nt + (g & 1)
                for h in range(2):
                    for e in range(2):
                        k = 16 * s + 8 * h + 2 * t + e
                        b_frag[lane][h][e] = w[k, co] if k < 27 else 0.0
            acc[nt] = mma_16816(a_frag, b_frag, acc[nt])
    # what the lanes store: lane (g, t) -> channels 8t .. 8t+7 of voxels 16j+g and 16j+g+8
    got = np.zeros((16, C0))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for nt in range(4):
            got[g, 8 * t + 2 * nt: 8 * t + 2 * nt + 2] = acc[nt][lane][0:2]
            got[g + 8, 8 * t + 2 * nt: 8 * t + 2 * nt + 2] = acc[nt][lane][2:4]
    ref = np.zeros((16, C0))
    for v in range(16):
        x = 16 * j + v                               # tile-local x; window column x + kw (column 0 = x - 1)
        for tap in range(27):
            kd, kh, kw = tap // 9, (tap // 3) % 3, tap % 3
            ref[v] += win[kd, ly + kh, x + kw] * w[tap]
        ref[v] += bias
    assert np.abs(got - ref).max() < 1e-12
