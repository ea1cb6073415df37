"""The one witness of the hop convention that the reference tree itself holds, tied to the oracle.

Upstream QUDA's Dslash is not in /root/reference, but the plug-in's own conserved-current kernel body
lib/code_pieces/fixSinkContractions_noether_core.h:100-137 spells out the four hop blocks of the Wilson operator,

    - S(x)      (1 - g_d)~ U_d(x)        F(x+d)      + S(x)      (1 + g_d)~ U_d(x-d)^dag  F(x-d)
    + S(x+d)    (1 + g_d)~ U_d(x)^dag    F(x)        - S(x-d)    (1 - g_d)~ U_d(x-d)      F(x)          (times 1/4, summed over x),

with (1 -+ g_d)~ the reference's own tables gammas_tm_base.h:148-171 (checked here to be exactly 1 -+ gamma_d of the oracle's gamma matrices).
The kernel body is compiled for the CPU from where it lies (oracle/_ref/libqkxtm_ref.so).  Feeding it rank-1 "propagators" S = conj(chi) and
F = psi with chi supported on ONE site and psi on ONE neighbouring site isolates a single block, which must equal chi^dag applied to
orc_dslash(psi) at that site:   forward   -1/2 chi(x)^dag [(1 - g_d) U_d(x) psi(x+d)],   backward  +1/2 chi(x)^dag [(1 + g_d) U_d(x-d)^dag psi(x-d)].
That pins, per direction and orientation: the projector sign, U versus U^dag, WHICH site's link is used, the neighbour index (through
the even-odd site map), the colour index order of the SU(3) multiply, and the link that carries the anti-periodic sign."""
import os

import numpy as np
import pytest

import lattice_util as lu
from oracle import ref as R
from oracle.oracle import Oracle, gamma_ukqcd

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libqkxtm_ref.so not built (needs /root/reference at build time)")

X = (4, 6, 4, 8)          # unequal extents, all >= 4: the eight neighbours of a site are distinct
V = int(np.prod(X))
PARTICLE, PARTFLAG = 0, 1


def lex(c):
    return c[0] + X[0] * (c[1] + X[1] * (c[2] + X[2] * c[3]))


def shift(c, d, s):
    c = list(c); c[d] = (c[d] + s) % X[d]; return tuple(c)


@pytest.fixture(scope="module")
def S():
    o = Oracle(X)
    r = R.Ref(X)
    gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=-1)                       # oracle layout: QDP even-odd, boundary sign folded in
    U_lex = lu.r2c(np.stack([lu.spinor_lex_from_eo(gauge[mu], X) for mu in range(4)]))   # [4][x_lex][3][3]
    gq = lu.c2r(np.ascontiguousarray(np.transpose(U_lex, (0, 2, 3, 1))))          # QKXTM layout [dir][c1][c2][x][re,im]
    gam = gamma_ukqcd()
    g5 = o.gamma5()
    perm = lu.eo_from_lex(X)
    inv = np.empty(V, dtype=np.int64); inv[perm] = np.arange(V)
    # the reference's own tables for the hop projectors (lib/code_pieces/gammas_tm_base.h:148-171, get_Operator flags 16+d and 20+d) are
    # exactly 1 + gamma_d and 1 - gamma_d of the oracle's gamma matrices (no twisted rotation on these entries)
    import ctypes as C
    for d in range(4):
        for f, sg in ((16 + d, +1), (20 + d, -1)):
            a = np.zeros(32)
            r.L.qref_get_operator(a.ctypes.data_as(C.POINTER(C.c_double)), f, PARTICLE, PARTFLAG)
            assert np.abs((a[0::2] + 1j * a[1::2]).reshape(4, 4) - (np.eye(4) + sg * gam[d])).max() < 1e-15, (d, sg)
    W = np.eye(4, dtype=np.complex128)
    return dict(o=o, r=r, gauge=gauge, gq=gq, W=W, inv=inv, U_lex=U_lex, gam=gam)


def reference_block(S, chi, xsite, psi, ysite):
    """sum over all sites of the reference kernel's accum[dir] (zero momentum, all time slices), for S = conj(W chi) at xsite and
    F = W^dag psi at ysite, both with source spin-colour (0, 0)"""
    W = S["W"]
    seq = np.zeros((4, 4, 3, 3, V), dtype=np.complex128)
    fwd = np.zeros((4, 4, 3, 3, V), dtype=np.complex128)
    seq[:, 0, :, 0, lex(xsite)] = np.conj(W @ chi)                  # S[ku][pu=0]^{c1, c2=0}(x) = conj((W chi)_{ku c1})
    fwd[:, 0, :, 0, lex(ysite)] = W.conj().T @ psi                  # F[lu][pu=0]^{c3, c2=0}(y) = (W^dag psi)_{lu c3}
    n, _ = S["r"].fixsink_derivative(lu.c2r(fwd), lu.c2r(seq), S["gq"], PARTICLE, PARTFLAG, [[0, 0, 0]], (0, 0, 0))
    return lu.r2c(n)[:, 0, :].sum(axis=0)                           # [dir]


def oracle_hop_at(S, psi, ysite, xsite, dagger=0):
    """(D psi_delta)(x) from orc_dslash for psi supported on the single site y"""
    o, inv = S["o"], S["inv"]
    py = sum(ysite) & 1
    field = np.zeros((o.Vh, 4, 3, 2))
    field[inv[lex(ysite)] - py * o.Vh] = lu.c2r(psi)
    out = o.dslash(S["gauge"], field, 1 - py, dagger)
    assert (sum(xsite) & 1) == 1 - py
    return lu.r2c(out[inv[lex(xsite)] - (1 - py) * o.Vh])           # [4][3]


@pytest.mark.parametrize("ysite", [(0, 0, 0, 0), (1, 4, 2, 5), (3, 5, 3, 7)])
@pytest.mark.parametrize("d", [0, 1, 2, 3])
def test_each_hop_block_of_the_oracle_equals_the_reference_kernels(S, d, ysite):
    rng = np.random.default_rng(1000 + 10 * d + sum(ysite))
    psi = rng.normal(size=(4, 3)) + 1j * rng.normal(size=(4, 3))
    chi = rng.normal(size=(4, 3)) + 1j * rng.normal(size=(4, 3))
    # forward block: output site x = y - d uses U_d(x) and (1 - g_d) on psi(x + d)
    x = shift(ysite, d, -1)
    ref_f = reference_block(S, chi, x, psi, ysite)
    want_f = -0.5 * np.vdot(chi, oracle_hop_at(S, psi, ysite, x))
    scale = np.linalg.norm(chi) * np.linalg.norm(psi)
    assert abs(ref_f[d] - want_f) < 1e-13 * scale, (d, ysite, ref_f[d], want_f)
    assert all(abs(ref_f[k]) < 1e-14 * scale for k in range(4) if k != d)
    # backward block: output site x = y + d uses U_d(y)^dag = U_d(x - d)^dag and (1 + g_d) on psi(x - d)
    x = shift(ysite, d, +1)
    ref_b = reference_block(S, chi, x, psi, ysite)
    want_b = +0.5 * np.vdot(chi, oracle_hop_at(S, psi, ysite, x))
    assert abs(ref_b[d] - want_b) < 1e-13 * scale, (d, ysite, ref_b[d], want_b)
    # and the blocks are what the formula says (independent numpy restatement on the lexicographic field)
    U = S["U_lex"]
    g = S["gam"][d]
    blk_f = (np.eye(4) - g) @ psi @ U[d, lex(shift(ysite, d, -1))].T            # (1 - g_d) (x) U_d(x) on psi(y)
    blk_b = (np.eye(4) + g) @ psi @ U[d, lex(ysite)].conj()                      # (1 + g_d) (x) U_d(y)^dag on psi(y)
    assert np.abs(oracle_hop_at(S, psi, ysite, shift(ysite, d, -1)) - blk_f).max() < 1e-13
    assert np.abs(oracle_hop_at(S, psi, ysite, shift(ysite, d, +1)) - blk_b).max() < 1e-13


def test_dagger_swaps_the_projectors_of_the_pinned_blocks(S):
    """D^dag has (1 + g_d) on the forward and (1 - g_d) on the backward hop: the same links and neighbours as the pinned blocks"""
    rng = np.random.default_rng(7)
    psi = rng.normal(size=(4, 3)) + 1j * rng.normal(size=(4, 3))
    y = (2, 3, 1, 7)
    U = S["U_lex"]
    for d in range(4):
        g = S["gam"][d]
        xf, xb = shift(y, d, -1), shift(y, d, +1)
        assert np.abs(oracle_hop_at(S, psi, y, xf, 1) - (np.eye(4) + g) @ psi @ U[d, lex(xf)].T).max() < 1e-13
        assert np.abs(oracle_hop_at(S, psi, y, xb, 1) - (np.eye(4) - g) @ psi @ U[d, lex(y)].conj()).max() < 1e-13


def test_time_boundary_link_carries_the_antiperiodic_sign(S):
    """y at t = 0: its backward neighbour in time sits on the last slice, whose U_t carries the folded-in -1 (qkxtm/QKXTM_util.cpp:698-705);
    the reference kernel and the oracle are given the same numbers, so the sign shows up identically in both"""
    U = S["U_lex"]
    raw = lu.r2c(np.stack([lu.spinor_lex_from_eo(lu.random_gauge_qdp(X, seed=137, t_boundary=+1)[mu], X) for mu in range(4)]))
    last = lex((1, 2, 3, X[3] - 1))
    assert np.allclose(U[3, last], -raw[3, last]) and np.allclose(U[3, lex((1, 2, 3, 2))], raw[3, lex((1, 2, 3, 2))])
