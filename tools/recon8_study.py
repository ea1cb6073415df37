#!/usr/bin/env python
"""CPU study for the next byte-reduction step of the Dslash: an 8-real link format that needs NO
trigonometric functions to unpack.  Stored per link: a1 = U01, a2 = U02, b1 = U10 (6 reals) and tan(arg(U00)/4), tan(arg(U20)/4);
|U00|^2 = 1 - |a1|^2 - |a2|^2, |U20|^2 = |a1|^2 + |a2|^2 - |b1|^2, the phases come back through the rational half-angle formulas
(1 + t^2 in [1, 2]: no singular point), U11 / U12 from row orthogonality and the determinant condition, the third row as in the
12-real format (with the same boundary sign).  Prints the reconstruction error over random SU(3) links (the generator of the parity
tests), incl. links carrying the anti-periodic sign."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lattice_util as lu  # noqa: E402


def pack(U):
    return U[:, 0, 1], U[:, 0, 2], U[:, 1, 0], np.tan(np.angle(U[:, 0, 0]) / 4), np.tan(np.angle(U[:, 2, 0]) / 4)


def phase(t):
    d = 1 / (1 + t * t); c2 = (1 - t * t) * d; s2 = 2 * t * d          # cos, sin of theta / 2
    return (c2 * c2 - s2 * s2) + 1j * (2 * s2 * c2)


def unpack(a1, a2, b1, t00, t20, sign):
    rs = np.abs(a1) ** 2 + np.abs(a2) ** 2
    u00 = np.sqrt(np.maximum(1 - rs, 0)) * phase(t00)
    u20 = np.sqrt(np.maximum(rs - np.abs(b1) ** 2, 0)) * phase(t20)
    rhs1 = -b1 * np.conj(u00)                                            # conj(a1) U11 + conj(a2) U12 = -U10 conj(U00)
    rhs2 = sign * np.conj(u20)                                           # -a2 U11 + a1 U12 = sign conj(U20)
    u11 = (rhs1 * a1 - np.conj(a2) * rhs2) / rs
    u12 = (np.conj(a1) * rhs2 + a2 * rhs1) / rs
    r = np.stack([u00, a1, a2], 1); s = np.stack([b1, u11, u12], 1)
    return np.stack([r, s, sign[:, None] * np.conj(np.cross(r, s))], 1)


if __name__ == "__main__":
    X = (16, 16, 16, 16)
    U = lu.random_su3_lex(X, seed=137).reshape(-1, 3, 3)
    sign = np.ones(len(U)); sign[::7] = -1                               # some links carry the anti-periodic boundary sign
    U = U * sign[:, None, None]
    V = unpack(*pack(U), sign)
    err = np.abs(V - U).reshape(len(U), -1).max(1)
    rs = np.abs(U[:, 0, 1]) ** 2 + np.abs(U[:, 0, 2]) ** 2
    print("links %d  max |dU| %.3e  median %.3e  99.9%% %.3e  relative L2 %.3e  (min row sum %.3e, min |U20| %.3e)"
          % (len(U), err.max(), np.median(err), np.quantile(err, 0.999), np.linalg.norm(V - U) / np.linalg.norm(U), rs.min(), np.abs(U[:, 2, 0]).min()))
