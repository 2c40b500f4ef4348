"""Lane-level CPU emulation of csrc/pm_conv.cu (test infrastructure only).

Mirrors the kernel step by step -- halo staging with the bank-conflict-free pixel stride, zero padding / zero
stuffing, 64-bit A-fragment addresses (k slots t, t+4 <- channels 2t, 2t+1), host-packed B fragments, the m16n8k8 lane <-> matrix-element mapping of
`mma.sync` and the epilogue's (pixel, channel) ownership -- in exact fp32/fp64 arithmetic, so that the CPU build
box can check the index math and the host-side filter packing against torch's conv2d without a GPU.
"""
from __future__ import annotations

import numpy as np


def round_kcin(cin: int) -> int:
    return 8 if cin <= 8 else (16 if cin <= 16 else (32 if cin <= 32 else 64))


def round_nt(cout: int) -> int:
    nt = (cout + 7) // 8
    return nt if nt <= 4 else 8


def pixel_stride(kcin: int, S: int) -> int:
    ps = kcin
    while (S * ps) % 32 not in (8, 24):
        ps += 4
    return ps


def effective_mt(nt: int, mt: int) -> int:
    if nt <= 2 and mt == 4:
        return 4
    if nt <= 4 and mt >= 2:
        return 2
    return 1


def emulate_conv(x_nhwc: np.ndarray, frag: np.ndarray, bias, cout: int, ks: int, S: int, pad: int, dil: int,
                 relu: bool, stuff: bool, mt: int, ycs: int = 0, yco: int = 0, y=None) -> np.ndarray:
    N, H, W, Cin = x_nhwc.shape
    kc, nt = round_kcin(Cin), round_nt(cout)
    KK = kc // 8
    mt = effective_mt(nt, mt)
    Hv, Wv = (2 * H, 2 * W) if stuff else (H, W)
    Ho = (Hv + 2 * pad - dil * (ks - 1) - 1) // S + 1
    Wo = (Wv + 2 * pad - dil * (ks - 1) - 1) // S + 1
    rows = 4 * mt
    tiles_x, tiles_y = (Wo + 15) // 16, (Ho + rows - 1) // rows
    rw = 15 * S + dil * (ks - 1) + 1
    rh = (rows - 1) * S + dil * (ks - 1) + 1
    ps = pixel_stride(kc, S)
    ycs = ycs or cout
    if y is None:
        y = np.full((N, Ho, Wo, ycs), np.nan, dtype=np.float64)
    frag = np.asarray(frag, dtype=np.float64).reshape(ks * ks, KK, nt, 32, -1)
    if frag.shape[-1] == 4:  # 3xTF32 packing (b0_hi, b1_hi, b0_lo, b1_lo): hi + lo restores the weight to ~2^-22
        frag = frag[..., :2] + frag[..., 2:]
    lanes = np.arange(32)
    g, t = lanes >> 2, lanes & 3
    banks_ok = True
    for n in range(N):
        for ty in range(tiles_y):
            for tx in range(tiles_x):
                ox0, oy0 = tx * 16, ty * rows
                ix0, iy0 = ox0 * S - pad, oy0 * S - pad
                s_in = np.zeros(rh * rw * ps, dtype=np.float64)
                for pix in range(rh * rw):
                    ry, rx = divmod(pix, rw)
                    iy, ix = iy0 + ry, ix0 + rx
                    inside = 0 <= iy < Hv and 0 <= ix < Wv
                    if stuff:
                        inside = inside and ((iy | ix) & 1) == 0
                        iy >>= 1
                        ix >>= 1
                    if inside:
                        s_in[pix * ps: pix * ps + Cin] = x_nhwc[n, iy, ix, :]
                for warp in range(4):
                    for m in range(mt):
                        orow = warp * mt + m
                        acc = np.zeros((nt, 32, 4), dtype=np.float64)
                        for ky in range(ks):
                            for kx in range(ks):
                                tap = ky * ks + kx
                                for kk in range(KK):
                                    base = ((orow * S + ky * dil) * rw + kx * dil) * ps + kk * 8
                                    i0 = base + g * S * ps + 2 * t   # 64-bit load: channels 2t, 2t+1 of pixel g
                                    i1 = i0 + 8 * S * ps
                                    for half in (slice(0, 16), slice(16, 32)):  # a 64-bit shared load is served per half-warp
                                        words = np.concatenate((i0[half], i0[half] + 1)) % 32
                                        banks_ok = banks_ok and len(set(words.tolist())) == 32
                                    A = np.zeros((16, 8))
                                    A[g, t] = s_in[i0]          # k slot t   <- channel 2t
                                    A[g + 8, t] = s_in[i1]
                                    A[g, t + 4] = s_in[i0 + 1]  # k slot t+4 <- channel 2t+1
                                    A[g + 8, t + 4] = s_in[i1 + 1]
                                    for j in range(nt):
                                        Bm = np.zeros((8, 8))
                                        Bm[t, g] = frag[tap, kk, j, :, 0]
                                        Bm[t + 4, g] = frag[tap, kk, j, :, 1]
                                        C = A @ Bm
                                        acc[j, :, 0] += C[g, 2 * t]
                                        acc[j, :, 1] += C[g, 2 * t + 1]
                                        acc[j, :, 2] += C[g + 8, 2 * t]
                                        acc[j, :, 3] += C[g + 8, 2 * t + 1]
                        oy = oy0 + orow
                        if oy >= Ho:
                            continue
                        for j in range(nt):
                            for lane in range(32):
                                co = j * 8 + 2 * t[lane]
                                for h in range(2):
                                    ox = ox0 + g[lane] + 8 * h
                                    if ox >= Wo:
                                        continue
                                    for e in range(2):
                                        if co + e >= cout:
                                            continue
                                        v = acc[j, lane, 2 * h + e] + (0.0 if bias is None else float(bias[co + e]))
                                        if relu:
                                            v = max(v, 0.0)
                                        y[n, oy, ox, yco + co + e] = v
    emulate_conv.last_banks_ok = banks_ok
    return y
