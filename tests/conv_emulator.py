"""Bit-level software model of conv_igemm_kernel's data flow (TMA boxes with zero fill, 128-B swizzle, weight blocks,
tap stacking, accumulator bookkeeping).  It consumes the REAL packed weight bytes produced by
oai_pack_conv_weights(), so the CPU test tier validates the packer and the issue-loop arithmetic without a GPU.
Hardware semantics (descriptor encodings) are only checked by the GPU tier.
"""
import numpy as np

MODE_ROW_SHARED, MODE_PER_TAP, MODE_POINTWISE = 0, 1, 2


def _xor(r, nchunks):
    """16-byte chunk XOR pattern of a row: SWIZZLE_128B (8 chunks): r % 8; SWIZZLE_64B (4 chunks): (r // 2) % 4."""
    return r % 8 if nchunks == 8 else (r // 2) % 4


def _swizzled_store(rows16):
    """rows16: [nrows, 64 | 32] uint16 -> smem image (uint16 view), chunks XOR-swizzled by the row index."""
    n, elems = rows16.shape
    nch = elems // 8
    out = np.zeros((n, nch, 8), dtype=np.uint16)
    r = np.arange(n)
    src = rows16.reshape(n, nch, 8)
    for j in range(nch):
        out[r, j ^ _xor(r, nch), :] = src[:, j, :]
    return out.reshape(n * elems)


def _read_operand(img16, start_byte, nrows, elems=64):
    """UMMA K-major swizzled read: row i at start + i*row_bytes, XOR pattern from the absolute row index."""
    rb = elems * 2
    assert start_byte % rb == 0
    nch = elems // 8
    r_abs = start_byte // rb + np.arange(nrows)
    img = img16.reshape(-1, nch, 8)
    out = np.zeros((nrows, nch, 8), dtype=np.uint16)
    for j in range(nch):
        out[:, j, :] = img[r_abs, j ^ _xor(r_abs, nch), :]
    return out.reshape(nrows, elems)


def _box(src, n, dp, ch, cw, cc, bh, bw, elems=64):
    """TMA tiled load with zero fill: src [NT,D,H,W,C] uint16 -> [bh*bw, elems]."""
    NT, D, H, W, C = src.shape
    out = np.zeros((bh, bw, elems), dtype=np.uint16)
    if 0 <= dp < D:
        for i in range(bh):
            h = ch + i
            if not 0 <= h < H:
                continue
            w_lo, w_hi = max(cw, 0), min(cw + bw, W)
            c_hi = min(cc + elems, C)
            if w_lo < w_hi and cc < c_hi:
                out[i, w_lo - cw:w_hi - cw, :c_hi - cc] = src[n, dp, h, w_lo:w_hi, cc:c_hi]
    return out.reshape(bh * bw, elems)


def build_chunks(c0, c1, terms):
    """(source, first channel) of every K chunk, in the order api_conv.cu::build_chunks lists them."""
    chunks = []
    for src, c in ((0, c0), (1, c1)):
        if c == 0:
            continue
        wide = 2 * c if terms >= 2 else c
        chunks += [(src, j * 64) for j in range((wide + 63) // 64)]
        if terms == 3:
            chunks += [(src, j * 64) for j in range((c + 63) // 64)]
    return chunks


def emulate(x0, x1, wpack, bias, plan, cout, relu=True, k16_steps=4, fp="f16", region=None, terms=1, split=False):
    """x0/x1: [NT,D,H,W,C] float16 arrays (x1 may be None; [.., 2C] = [hi | lo] planes when split); wpack: uint8 array;
    returns float32 [NT,D,H,W,cout]."""
    f16 = np.float16
    NT, D, H, W, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[-1]
    if split:
        c0, c1 = c0 // 2, c1 // 2
    chunks = build_chunks(c0, c1, terms)
    mode, kpb, R, nhalf, cph, nblk, wbytes = (plan[k] for k in ("mode", "kd_per_block", "R", "nhalf",
                                                                  "cout_per_half", "nblk", "wblock_bytes"))
    TW = min(W, 128)
    TH = 128 // TW
    elems = plan.get("row_bytes", 128) // 2
    rb = elems * 2
    xs = (x0.view(np.uint16), None if x1 is None else x1.view(np.uint16))
    w16 = np.frombuffer(wpack.tobytes(), dtype=np.uint16)
    out = np.zeros((NT, D, H, W, cout), dtype=np.float32)
    nkw = 3 if mode == MODE_ROW_SHARED else 1
    d_lo, d_cnt, h_lo, h_cnt = region if region is not None else (0, D, 0, H)
    hp_lo, hp_hi = h_lo // TH, (h_lo + h_cnt + TH - 1) // TH
    for n in range(NT):
        for d0 in range(d_lo, d_lo + d_cnt, R):
            for h0 in range(hp_lo * TH, hp_hi * TH, TH):
                for w0 in range(0, W, TW):
                    rv = min(R, d_lo + d_cnt - d0)  # the last d-group of a region may be partial
                    for nh in range(nhalf):
                        acc = np.zeros((R, 128, cph), dtype=np.float32)
                        touched = [False] * R
                        for b in range(nblk):
                            if mode == MODE_ROW_SHARED:
                                c, kh, kw0, kdlo, nkd = b // 3, b % 3, 0, 0, 3
                            elif mode == MODE_PER_TAP and kpb == 3:
                                c, r = b // 9, b % 9
                                kh, kw0, kdlo, nkd = r // 3, r % 3, 0, 3
                            elif mode == MODE_PER_TAP:
                                c, r = b // 27, b % 27
                                kh, kw0, kdlo, nkd = r // 9, (r // 3) % 3, r % 3, 1
                            else:
                                c, kh, kw0, kdlo, nkd = b, 1, 1, 1, 1
                            blk = w16[((nh * nblk + b) * wbytes) // 2:((nh * nblk + b + 1) * wbytes) // 2]
                            src, cc = xs[chunks[c][0]], chunks[c][1]
                            kdhi = kdlo + nkd - 1
                            dlo, dhi = max(0, d0 + kdlo - 1), min(D - 1, d0 + rv - 1 + kdhi - 1)
                            if mode == MODE_ROW_SHARED:
                                cw, bw, bh = w0 - 1, 130, 1
                            elif mode == MODE_PER_TAP:
                                cw, bw, bh = w0 + kw0 - 1, TW, TH
                            else:
                                cw, bw, bh = w0, TW, TH
                            ch = h0 + kh - 1
                            for dp in range(dlo, dhi + 1):
                                stage = _swizzled_store(_box(src, n, dp, ch, cw, cc, bh, bw, elems))
                                a_first = dp - kdhi + 1 - d0
                                ti_lo, ti_hi = max(0, -a_first), min(nkd - 1, rv - 1 - a_first)
                                for kw in range(nkw):
                                    A = _read_operand(stage, kw * rb, 128, elems).view(f16).astype(np.float32)
                                    A[:, k16_steps * 16:] = 0
                                    ti = ti_lo
                                    while ti <= ti_hi:
                                        a0 = a_first + ti
                                        f = touched[a0]
                                        ln = 1
                                        while ti + ln <= ti_hi and touched[a0 + ln] == f and (ln + 1) * cph <= 256:
                                            ln += 1
                                        Bm = _read_operand(blk, (kw * nkd + ti) * cph * rb, ln * cph, elems)
                                        Bm = Bm.view(f16).astype(np.float32)
                                        prod = A @ Bm.T  # [128, ln*cph]
                                        for j in range(ln):
                                            if f:
                                                acc[a0 + j] += prod[:, j * cph:(j + 1) * cph]
                                            else:
                                                acc[a0 + j] = prod[:, j * cph:(j + 1) * cph]
                                            touched[a0 + j] = True
                                        ti += ln
                        for a in range(rv):
                            v = acc[a] + bias[nh * cph:(nh + 1) * cph][None, :]
                            if relu:
                                v = np.maximum(v, 0)
                            out[n, d0 + a, h0:h0 + TH, w0:w0 + TW, nh * cph:(nh + 1) * cph] = v.reshape(TH, TW, cph)
    return out
