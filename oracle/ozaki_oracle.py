"""CPU restatement of the integer arithmetic of csrc/syrk_i8.cu -- TEST INFRASTRUCTURE ONLY.

This is not a reference-side algorithm (the reference's SYRK lives inside Ceres' Schur eliminator, in plain FP64);
it restates OUR tensor-core formulation so that its arithmetic can be checked without a GPU:
  * column scales 2^e >= max|z|, x = z 2^-e rounded to B = 8s-2 fractional bits;
  * balanced base-256 digits from one add: bytes of (X + 0x80..80) xor 0x80;
  * exact integer pair products C_t = sum_{p+q=t} D_p^T D_q for t <= s+1, recombined in float64.
tests/test_ozaki_oracle.py checks digits, exactness and the error bound against numpy float64 / exact integers.
"""
from __future__ import annotations

import numpy as np


def column_exponents(Z):
    """e_d with |Z[:, d]| 2^-e_d < 1 (ilogb(max) + 1); 0 for all-zero columns."""
    m = np.abs(Z).max(axis=0)
    e = np.zeros(Z.shape[1], dtype=np.int64)
    nz = m > 0
    e[nz] = np.frexp(m[nz])[1]            # m = f 2^e, f in [0.5, 1)  ->  m 2^-e < 1
    return e


def slices(Z, s):
    """int8 digit matrices D[p] (p = 0 most significant) and exponents e: Z ~ 2^(e - B) sum_p D[p] 256^(s-1-p)."""
    B = 8 * s - 2
    e = column_exponents(Z)
    X = np.rint(np.ldexp(Z, (B - e)[None, :])).astype(np.int64)
    bias = int.from_bytes(b"\x80" * s, "little")
    Y = (X + bias) ^ bias
    D = np.empty((s,) + Z.shape, dtype=np.int8)
    for p in range(s):
        j = s - 1 - p
        D[p] = ((Y >> (8 * j)) & 255).astype(np.uint8).view(np.int8)
    return D, e, X


def syrk(Z, s, max_order=None):
    """Z^T Z through the sliced integer products, recombined like the kernel's epilogue (float64)."""
    D, e, _ = slices(Z, s)
    B = 8 * s - 2
    tmax = s + 1 if max_order is None else max_order
    n = Z.shape[1]
    out = np.zeros((n, n))
    for t in range(tmax, 1, -1):                       # least significant order first
        C = np.zeros((n, n), dtype=np.int64)
        for p in range(1, s + 1):
            q = t - p
            if 1 <= q <= s:
                C += D[p - 1].astype(np.int64).T @ D[q - 1].astype(np.int64)
        assert np.abs(C).max() < 2 ** 31, "int32 accumulator would overflow"
        out += np.ldexp(C.astype(np.float64), 8 * (2 * s - t) - 2 * B)
    return np.ldexp(out, (e[:, None] + e[None, :]))
