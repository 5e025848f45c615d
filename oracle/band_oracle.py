"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the band detection that `vgg_ba_solve` runs for sequential (video) problems
(vggsfm_b200/csrc/ba_solve.cu, compute_band_hint): from the visibility mask [S, N] to
  * the 64-row k-block range per 128-column row block of the Schur operand Zt (rows 3n+c, columns s*dc+i, then the
    shared intrinsics) outside which that block is exactly zero,
  * the block structure (band end per block column + arrow block) the factorisation assumes,
  * the per-k-block row range `backsub` walks and the per-frame-group track range `ba_blocks` keeps.
The reference has no counterpart (pycolmap/Ceres exploit the same sparsity inside SPARSE_SCHUR, video_runner.py:508);
tests/test_band_oracle.py checks these tables against the brute-force sparsity pattern of random banded masks, i.e. that
nothing non-zero can fall outside what the kernels visit.
"""
from __future__ import annotations

import numpy as np


def frame_ranges(mask):
    """first / last visible point of every frame ((N, -1) when the frame sees nothing)."""
    S, N = mask.shape
    fr = np.zeros((S, 2), dtype=np.int64)
    for s in range(S):
        idx = np.nonzero(mask[s])[0]
        fr[s] = (idx[0], idx[-1]) if idx.size else (N, -1)
    return fr


def band_tables(mask, dc, ns):
    """Returns dict(rb_range [nb,2], end_blk [nb], arrow_blk, kb_rows [KB,2], fg_tracks [ngroups,2]) with the formulas of
    compute_band_hint (no `is it worth it` threshold)."""
    S, N = mask.shape
    D = S * dc + ns
    Dpad = (D + 2 + 127) // 128 * 128
    Kpad = (3 * N + 63) // 64 * 64
    nb, KB = Dpad // 128, Kpad // 64
    fr = frame_ranges(mask)
    rg = np.zeros((nb, 2), dtype=np.int64)
    for rb in range(nb):
        d0, d1 = rb * 128, min(D, rb * 128 + 128) - 1
        lo, hi = KB, 0
        if d1 >= d0:
            if d1 >= S * dc:
                lo, hi = 0, KB
            else:
                for f in range(d0 // dc, d1 // dc + 1):
                    if fr[f, 1] < 0:
                        continue
                    lo = min(lo, 3 * fr[f, 0] // 64)
                    hi = max(hi, (3 * fr[f, 1] + 2) // 64 + 1)
        if hi <= lo:
            lo = hi = 0
        rg[rb] = (lo, min(hi, KB))
    arrow = (S * dc) // 128
    end = np.zeros(nb, dtype=np.int64)
    prev = 0
    for b in range(nb):
        if b < arrow:
            e = b
            for i in range(b + 1, arrow):
                if min(rg[i, 1], rg[b, 1]) > max(rg[i, 0], rg[b, 0]):
                    e = i
            e = max(e + 1, min(b + 2, arrow))
            e = min(max(e, prev), arrow)
        else:
            e = nb
        end[b] = e
        prev = e
    kb_rows = np.zeros((KB, 2), dtype=np.int64)
    for kb in range(KB):
        rbs = [rb for rb in range(arrow) if rg[rb, 0] <= kb < rg[rb, 1]]
        if rbs:
            kb_rows[kb] = (rbs[0] * 128, (rbs[-1] + 1) * 128)
    ng = (S + 31) // 32
    fg = np.zeros((ng, 2), dtype=np.int64)
    for g in range(ng):
        lo, hi = N, 0
        for f in range(32 * g, min(S, 32 * g + 32)):
            if fr[f, 1] < 0:
                continue
            lo, hi = min(lo, fr[f, 0]), max(hi, fr[f, 1] + 1)
        fg[g] = (lo, hi) if hi > lo else (0, 0)
    return {"rb_range": rg, "end_blk": end, "arrow_blk": arrow, "kb_rows": kb_rows, "fg_tracks": fg, "nb": nb, "KB": KB,
            "D": D, "Dpad": Dpad}


def brute_force_pattern(mask, dc, ns):
    """Boolean sparsity of Zt [3N, D] (point n touches the dc columns of every frame that sees it + the shared columns)
    and of the reduced system S = H_cc - Zt^T Zt [D, D] (+ the bordered right-hand-side row)."""
    S, N = mask.shape
    D = S * dc + ns
    Z = np.zeros((3 * N, D), dtype=bool)
    cam = np.repeat(mask.T, dc, axis=1)                      # [N, S*dc]
    Z[:, :S * dc] = np.repeat(cam, 3, axis=0)
    Z[:, S * dc:] = np.repeat(mask.any(0)[:, None], 3, axis=0)
    Zf = Z.astype(np.float32)                                # BLAS matmul; counts stay far below 2^24
    ZtZ = (Zf.T @ Zf) > 0.5
    return Z, ZtZ | np.eye(D, dtype=bool), ZtZ
