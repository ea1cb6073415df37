"""Pins for the CPU oracle (oracle/tm_oracle.c).  The reference holds no golden vectors for this
path ("parity unpinned"), so the oracle is pinned by algebraic identities,
by an independent dense numpy restatement on the non-checkerboarded lattice, and by frozen
checksums (tests/golden/) that guard against silent drift."""
import json
import os

import numpy as np
import pytest

import lattice_util as lu
from oracle.oracle import Oracle, gamma_ukqcd, gamma_degrand_rossi

X4 = (4, 4, 4, 4)
XA = (4, 6, 4, 8)   # anisotropic extents catch x/y/z/t mix-ups
KAPPA = 1.0 / (2.0 * (4.0 + 0.1))   # default mass 0.1 (qkxtm/Calc_Loops.cpp:382-388)
MU = 0.1


def full_eo_to_lexc(orc_full, X):
    """[even|odd][4][3][2] -> complex [V][4][3] lexicographic"""
    return lu.r2c(lu.spinor_lex_from_eo(orc_full, X))


def test_gamma_algebra_and_gamma5():
    g = gamma_ukqcd()
    for mu in range(4):
        assert np.allclose(g[mu], g[mu].conj().T)
        for nu in range(4):
            assert np.allclose(g[mu] @ g[nu] + g[nu] @ g[mu], 2 * np.eye(4) * (mu == nu))
    o = Oracle(X4)
    g5 = o.gamma5()
    swap = np.zeros((4, 4)); swap[0, 2] = swap[1, 3] = swap[2, 0] = swap[3, 1] = 1   # apply_gamma5_vector_core.h:1-16
    assert np.allclose(g5, swap)
    # 1 -+ gamma_mu as tabulated in gammas_tm_base.h:148-171 (spot values)
    assert np.allclose((np.eye(4) + g[3]), np.diag([2, 2, 0, 0]))
    assert np.allclose((np.eye(4) - g[3]), np.diag([0, 0, 2, 2]))
    assert np.isclose((np.eye(4) + g[0])[0, 3], 1j) and np.isclose((np.eye(4) - g[1])[3, 0], -1)


def test_index_helpers_match_even_odd_order():
    o = Oracle(XA)
    perm = lu.eo_from_lex(XA)
    Vh = o.Vh
    for odd in (0, 1):
        for i in (0, 1, 5, Vh // 3, Vh - 1):
            assert o.full_index(i, odd) == perm[odd * Vh + i]
    x, y, z, t = lu.coords_lex(XA)
    for odd in (0, 1):
        i = Vh // 2 + 3
        Y = perm[odd * Vh + i]
        for (d4, d3, d2, d1) in [(0, 0, 0, 1), (0, 0, 0, -1), (0, 0, 1, 0), (0, -1, 0, 0), (1, 0, 0, 0), (-1, 0, 0, 0)]:
            nb = o.neighbor_index(i, odd, d4, d3, d2, d1)
            xn = ((x[Y] + d1) % XA[0], (y[Y] + d2) % XA[1], (z[Y] + d3) % XA[2], (t[Y] + d4) % XA[3])
            Yn = xn[0] + XA[0] * (xn[1] + XA[1] * (xn[2] + XA[2] * xn[3]))
            assert nb == Yn // 2


@pytest.mark.parametrize("X", [X4, XA])
@pytest.mark.parametrize("dagger", [0, 1])
def test_full_operator_matches_dense_numpy(X, dagger):
    o = Oracle(X)
    U = lu.random_su3_lex(X, seed=137)
    gq = lu.gauge_qdp_from_lex(U, X, t_boundary=-1)
    Ubc = U.copy(); xs, ys, zs, ts = lu.coords_lex(X); Ubc[3, ts == X[3] - 1] *= -1
    psi = lu.gaussian_spinor_lex(X, seed=101)
    ref = lu.dense_mat(Ubc, lu.r2c(psi), X, gamma_ukqcd(), KAPPA, MU, dagger=bool(dagger))
    got = o.mat(gq, lu.spinor_eo_from_lex(psi, X), KAPPA, MU, dagger)
    assert lu.rel_l2(full_eo_to_lexc(got, X).view(np.float64), ref.view(np.float64)) < 1e-14


def test_hop_matches_dense_numpy_both_parities():
    X = XA; o = Oracle(X)
    U = lu.random_su3_lex(X, seed=5)
    gq = lu.gauge_qdp_from_lex(U, X, t_boundary=+1)
    psi = lu.gaussian_spinor_lex(X, seed=7)
    peo = lu.spinor_eo_from_lex(psi, X)
    for dagger in (0, 1):
        ref = lu.dense_hop(U, lu.r2c(psi), X, gamma_ukqcd(), dagger=bool(dagger))
        ref_eo = lu.c2r(ref)[lu.eo_from_lex(X)]
        out_e = o.dslash(gq, np.ascontiguousarray(peo[o.Vh:]), 0, dagger)
        out_o = o.dslash(gq, np.ascontiguousarray(peo[:o.Vh]), 1, dagger)
        assert lu.rel_l2(out_e, ref_eo[:o.Vh]) < 1e-14
        assert lu.rel_l2(out_o, ref_eo[o.Vh:]) < 1e-14


@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
def test_adjointness_matpc(matpc):
    o = Oracle(XA)
    gq = lu.random_gauge_qdp(XA)
    x = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(XA, seed=1), XA)[:o.Vh].copy()
    y = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(XA, seed=2), XA)[:o.Vh].copy()
    My = o.matpc(gq, y, KAPPA, MU, matpc, 0)
    Mdx = o.matpc(gq, x, KAPPA, MU, matpc, 1)
    lhs = o.cdot(x, My); rhs = o.cdot(Mdx, y)
    assert abs(lhs - rhs) / abs(lhs) < 1e-13


def test_adjointness_and_gamma5_hermiticity_full():
    o = Oracle(XA)
    gq = lu.random_gauge_qdp(XA)
    x = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(XA, seed=1), XA)
    y = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(XA, seed=2), XA)
    lhs = o.cdot(x, o.mat(gq, y, KAPPA, MU, 0)); rhs = o.cdot(o.mat(gq, x, KAPPA, MU, 1), y)
    assert abs(lhs - rhs) / abs(lhs) < 1e-13
    # gamma5 M(mu) gamma5 = M(-mu)^dag
    def g5(v):
        out = np.empty_like(v); out[:, 0] = v[:, 2]; out[:, 1] = v[:, 3]; out[:, 2] = v[:, 0]; out[:, 3] = v[:, 1]; return out
    a = g5(o.mat(gq, g5(y), KAPPA, MU, 0)); b = o.mat(gq, y, KAPPA, -MU, 1)
    assert lu.rel_l2(a, b) < 1e-14


def test_free_field_plane_wave_eigenvalue():
    """Unit gauge, periodic BC, mu = 0: M_full e^{ipx} chi = [1 - 2 kappa sum cos p + 2 i kappa sum g sin p] chi e^{ipx}"""
    X = XA; o = Oracle(X)
    gq = lu.unit_gauge_qdp(X, t_boundary=+1)
    n = (1, 2, 0, 3)
    p = [2 * np.pi * n[d] / X[d] for d in range(4)]
    x, y, z, t = lu.coords_lex(X)
    phase = np.exp(1j * (p[0] * x + p[1] * y + p[2] * z + p[3] * t))
    rng = np.random.default_rng(3)
    chi = rng.normal(size=(4, 3)) + 1j * rng.normal(size=(4, 3))
    psi = phase[:, None, None] * chi[None]
    g = gamma_ukqcd()
    Mp = (1 - 2 * KAPPA * sum(np.cos(p))) * np.eye(4) + 2j * KAPPA * sum(g[d] * np.sin(p[d]) for d in range(4))
    want = phase[:, None, None] * np.einsum("st,tc->sc", Mp, chi)[None]
    got = o.mat(gq, lu.spinor_eo_from_lex(lu.c2r(psi), X), KAPPA, 0.0, 0)
    assert lu.rel_l2(full_eo_to_lexc(got, X).view(np.float64), want.view(np.float64)) < 1e-13


def test_gauge_covariance():
    X = X4; o = Oracle(X)
    U = lu.random_su3_lex(X, seed=11)
    G = lu.random_su3_lex(X, seed=12)[0]           # one SU(3) rotation per site
    xs, ys, zs, ts = lu.coords_lex(X)
    V = len(xs)
    Urot = np.empty_like(U)
    for mu in range(4):
        c = [xs.copy(), ys.copy(), zs.copy(), ts.copy()]
        c[mu] = (c[mu] + 1) % X[mu]
        nb = c[0] + X[0] * (c[1] + X[1] * (c[2] + X[2] * c[3]))
        Urot[mu] = np.einsum("xab,xbc,xdc->xad", G, U[mu], np.conj(G[nb]))
    psi = lu.r2c(lu.gaussian_spinor_lex(X, seed=13))
    psir = np.einsum("xab,xsb->xsa", G, psi)
    def M(Ul, ps):
        gq = lu.gauge_qdp_from_lex(Ul, X, t_boundary=-1)
        return full_eo_to_lexc(o.mat(gq, lu.spinor_eo_from_lex(lu.c2r(ps), X), KAPPA, MU, 0), X)
    a = M(Urot, psir); b = np.einsum("xab,xsb->xsa", G, M(U, psi))
    assert lu.rel_l2(a.view(np.float64), b.view(np.float64)) < 1e-13


def test_change_of_basis_degrand_rossi():
    """derive S with S g_DR S^dag = g_UKQCD by group averaging, then M_UK (S psi) = S (M_DR psi)"""
    gU, gD = gamma_ukqcd(), gamma_degrand_rossi()
    for mu in range(4):
        for nu in range(4):
            assert np.allclose(gD[mu] @ gD[nu] + gD[nu] @ gD[mu], 2 * np.eye(4) * (mu == nu))
    rng = np.random.default_rng(0)
    Xr = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    S = np.zeros((4, 4), dtype=complex)
    for bits in range(16):
        a = np.eye(4, dtype=complex); b = np.eye(4, dtype=complex)
        for mu in range(4):
            if bits >> mu & 1:
                a = a @ gU[mu]; b = b @ gD[mu]
        S += a @ Xr @ np.linalg.inv(b)
    S = S / np.sqrt((S @ S.conj().T)[0, 0].real)
    assert np.allclose(S @ S.conj().T, np.eye(4))
    for mu in range(4):
        assert np.allclose(S @ gD[mu] @ S.conj().T, gU[mu])
    X = X4
    gq = lu.random_gauge_qdp(X)
    psi = lu.r2c(lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=21), X))[: np.prod(X) // 2]
    oD = Oracle(X, gamma=gD)
    # gamma5 is diagonal in DR (upstream host reference convention), off-diagonal in UKQCD
    assert np.allclose(np.abs(oD.gamma5()), np.eye(4))
    mD = lu.r2c(oD.matpc(gq, lu.c2r(psi), KAPPA, MU, 0, 0))
    oU = Oracle(X, gamma=gU)
    mU = lu.r2c(oU.matpc(gq, lu.c2r(np.einsum("st,xtc->xsc", S, psi)), KAPPA, MU, 0, 0))
    assert lu.rel_l2(mU.view(np.float64), np.einsum("st,xtc->xsc", S, mD).view(np.float64)) < 1e-13


@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
def test_schur_prepare_solve_reconstruct(matpc):
    """prepare -> M^dag -> CG on M^dag M -> reconstruct reproduces a full-lattice solve:
    M_full x = b  (call order of lib/qudaQKXTM_interface.cpp:2020-2041)"""
    X = X4; o = Oracle(X)
    gq = lu.random_gauge_qdp(X)
    b = lu.spinor_eo_from_lex(lu.z4_source_lex(X, seed=100), X)
    src = o.prepare(gq, b, KAPPA, MU, matpc)
    rhs = o.matpc(gq, src, KAPPA, MU, matpc, 1)
    xp, it, tr, _ = o.cg_mdagm(gq, rhs, KAPPA, MU, matpc, tol=1e-12, maxiter=2000)
    assert tr < 1e-11 and 0 < it < 2000
    p = matpc & 1
    x = np.zeros_like(b); x[p * o.Vh:(p + 1) * o.Vh] = xp
    x = o.reconstruct(gq, x, b, KAPPA, MU, matpc)
    res = o.mat(gq, x, KAPPA, MU, 0) - b
    assert np.linalg.norm(res) / np.linalg.norm(b) < 1e-10


def test_reconstruct12_and_boundary_sign():
    o = Oracle(X4)
    U = lu.random_su3_lex(X4, seed=31)[2, 17]
    m = lu.c2r(U).reshape(18)
    assert np.allclose(o.reconstruct12(m, 1.0), m, atol=1e-15)
    assert np.allclose(o.reconstruct12(-m, -1.0), -m, atol=1e-15)   # link with anti-periodic sign folded in
    assert np.allclose(U @ U.conj().T, np.eye(3), atol=1e-14) and np.isclose(np.linalg.det(U), 1.0)


def test_plaquette():
    o = Oracle(XA)
    assert abs(o.plaquette(lu.unit_gauge_qdp(XA)) - 1.0) < 1e-15
    U = lu.random_su3_lex(XA, seed=137)
    gq = lu.gauge_qdp_from_lex(U, XA, t_boundary=+1)
    shp = (XA[3], XA[2], XA[1], XA[0], 3, 3)
    Ug = [U[m].reshape(shp) for m in range(4)]
    tot = 0.0
    for mu in range(4):
        for nu in range(mu + 1, 4):
            a = np.einsum("...ab,...bc->...ac", Ug[mu], np.roll(Ug[nu], -1, axis=3 - mu))
            b = np.einsum("...ab,...bc->...ac", Ug[nu], np.roll(Ug[mu], -1, axis=3 - nu))
            tot += np.sum(a * np.conj(b)).real
    assert abs(o.plaquette(gq) - tot / (np.prod(XA) * 18)) < 1e-13


def test_cg_variants_agree():
    X = X4; o = Oracle(X)
    gq = lu.random_gauge_qdp(X)
    b = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)[:o.Vh].copy()
    x0, it0, tr0, h0 = o.cg_mdagm(gq, b, KAPPA, MU, 0, tol=1e-10, pr_beta=0)
    x1, it1, tr1, h1 = o.cg_mdagm(gq, b, KAPPA, MU, 0, tol=1e-10, pr_beta=1)
    assert abs(it0 - it1) <= 1 and tr0 < 2e-10 and tr1 < 2e-10
    assert lu.rel_l2(x0, x1) < 1e-9
    assert np.all(np.diff(np.log(h0)) < 5.0)


GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_checksums.json")


def oracle_checksums():
    out = {}
    for name, X in (("4x4x4x4", X4), ("8x8x8x16", (8, 8, 8, 16))):
        o = Oracle(X)
        gq = lu.random_gauge_qdp(X, seed=137)
        psi = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)[:o.Vh].copy()
        m = o.matpc(gq, psi, KAPPA, MU, 0, 0)
        flat = m.ravel()
        idx = [0, 1, 23, 24 * 7 + 5, flat.size // 3, flat.size // 2 + 11, flat.size - 25, flat.size - 1]
        out[name] = {"plaquette": o.plaquette(gq), "sum": float(flat.sum()), "sumsq": float((flat ** 2).sum()),
                     "samples": [float(flat[i]) for i in idx]}
    return out


def test_golden_checksums():
    """frozen M_pc psi checksums (generated by tests/golden/make_golden.py from this oracle)"""
    with open(GOLD) as f:
        gold = json.load(f)
    got = oracle_checksums()
    for name in gold:
        for key in ("plaquette", "sum", "sumsq"):
            assert abs(got[name][key] - gold[name][key]) <= 1e-11 * max(1.0, abs(gold[name][key])), (name, key)
        assert np.allclose(got[name]["samples"], gold[name]["samples"], rtol=1e-12, atol=1e-13)


def test_chebyshev_filter_restatement():
    """the oracle's polynomialOperator restatement (reference lib/qudaQKXTM_Deflation.cpp:997-1063) is a polynomial in A:
    on a diagonal operator it equals the scalar recurrence applied to each eigenvalue, p(0) = 1, |p| stays below its
    edge value inside the window [amin, amax] and grows monotonically below amin"""
    from oracle.oracle import poly_operator, poly_scalar, cheb_coefficients
    rng = np.random.default_rng(5)
    lam = np.concatenate([[0.0, 1e-3, 1e-2], rng.uniform(0.05, 4.0, 40)])
    x = rng.standard_normal(lam.size) + 1j * rng.standard_normal(lam.size)
    for deg in (0, 1, 2, 7, 40):
        y = poly_operator(lambda v: lam * v, x, deg, 0.05, 4.0)
        p = np.array([poly_scalar(l, deg, 0.05, 4.0) for l in lam])
        assert np.allclose(y, p * x, rtol=1e-12, atol=1e-14)
        assert abs(p[0] - 1.0) < 1e-12
        if deg >= 2:
            edge = abs(poly_scalar(0.05, deg, 0.05, 4.0))
            assert np.all(np.abs(p[3:]) <= edge * (1 + 1e-9))
            assert p[0] > p[1] > p[2] > edge
    assert len(cheb_coefficients(5, 0.05, 4.0)) == 5


# ---- twisted-clover --------------------------------------------------------------------------------
CSW = 1.57551           # the ETMC N_f = 2 clover ensembles' value; coeff = csw * kappa (qkxtm/MG_Bench.cpp:249)


def _clover_setup(X, seed=137):
    o = Oracle(X)
    U = lu.random_su3_lex(X, seed=seed)
    gq = lu.gauge_qdp_from_lex(U, X, t_boundary=-1)
    Ubc = U.copy(); xs, ys, zs, ts = lu.coords_lex(X); Ubc[3, ts == X[3] - 1] *= -1
    clov = o.clover_compute(gq, CSW * KAPPA)
    return o, U, Ubc, gq, clov


def test_clover_term_matches_dense_numpy_and_is_hermitian():
    X = XA
    o, U, Ubc, gq, clov = _clover_setup(X)
    V = o.V
    want = lu.dense_clover(Ubc, X, gamma_ukqcd(), CSW * KAPPA).reshape(V, 12, 12)
    got_eo = lu.r2c(clov)                                    # [even Vh | odd Vh][12][12]
    got = np.empty_like(got_eo); got[lu.eo_from_lex(X)] = got_eo
    assert np.abs(got - want).max() < 1e-14
    assert np.abs(got - np.conj(np.swapaxes(got, 1, 2))).max() < 1e-14      # C is hermitian
    # the anti-periodic sign on U_t(T-1) cancels in every leaf: same C as for the periodic field
    assert np.abs(want - lu.dense_clover(U, X, gamma_ukqcd(), CSW * KAPPA).reshape(V, 12, 12)).max() < 1e-14
    # unit gauge field: F = 0, C = 1
    unit = np.zeros_like(U); unit[:, :, range(3), range(3)] = 1.0
    assert np.abs(lu.dense_clover(unit, X, gamma_ukqcd(), 0.3).reshape(V, 12, 12) - np.eye(12)).max() == 0.0


@pytest.mark.parametrize("dagger", [0, 1])
def test_clover_full_operator_matches_dense_numpy(dagger):
    X = XA
    o, U, Ubc, gq, clov = _clover_setup(X)
    o.set_clover(clov)
    psi = lu.gaussian_spinor_lex(X, seed=101)
    ref = lu.dense_mat_clover(Ubc, lu.r2c(psi), X, gamma_ukqcd(), KAPPA, MU, CSW * KAPPA, dagger=bool(dagger))
    got = o.mat(gq, lu.spinor_eo_from_lex(psi, X), KAPPA, MU, dagger)
    o.set_clover(None)
    assert lu.rel_l2(full_eo_to_lexc(got, X).view(np.float64), ref.view(np.float64)) < 1e-14


def test_clover_site_operator_inverse_and_csw_zero_limit():
    X = X4
    o, U, Ubc, gq, clov = _clover_setup(X)
    psi = np.ascontiguousarray(lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=7), X)[: o.Vh])
    o.set_clover(clov)
    for parity in (0, 1):
        for dagger in (0, 1):
            y = o.site_A(psi, KAPPA, MU, parity, dagger, 0)
            assert lu.rel_l2(o.site_A(y, KAPPA, MU, parity, dagger, 1), psi) < 1e-14
    # csw = 0: C = 1, everything reduces to plain twisted mass
    o.set_clover(o.clover_compute(gq, 0.0))
    a = o.mdagm(gq, psi, KAPPA, MU, 0)
    o.set_clover(None)
    assert lu.rel_l2(a, o.mdagm(gq, psi, KAPPA, MU, 0)) < 1e-15


@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
def test_clover_adjointness_and_schur(matpc):
    """<x, M y> = <M^dag x, y> for the preconditioned twisted-clover operator, and prepare -> M^dag -> CG -> reconstruct
    solves the full system"""
    X = X4
    o, U, Ubc, gq, clov = _clover_setup(X)
    o.set_clover(clov)
    x = np.ascontiguousarray(lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=3), X)[: o.Vh])
    y = np.ascontiguousarray(lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=4), X)[: o.Vh])
    lhs = np.vdot(lu.r2c(x), lu.r2c(o.matpc(gq, y, KAPPA, MU, matpc, 0)))
    rhs = np.vdot(lu.r2c(o.matpc(gq, x, KAPPA, MU, matpc, 1)), lu.r2c(y))
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    b = lu.spinor_eo_from_lex(lu.z4_source_lex(X, seed=100), X)
    src = o.prepare(gq, b, KAPPA, MU, matpc)
    rhs_v = o.matpc(gq, src, KAPPA, MU, matpc, 1)
    xp, it, tr, _ = o.cg_mdagm(gq, rhs_v, KAPPA, MU, matpc, tol=1e-12, maxiter=2000)
    p = matpc & 1
    xf = np.zeros_like(b); xf[p * o.Vh:(p + 1) * o.Vh] = xp
    xf = o.reconstruct(gq, xf, b, KAPPA, MU, matpc)
    res = o.mat(gq, xf, KAPPA, MU, 0) - b
    o.set_clover(None)
    assert tr < 1e-11 and np.linalg.norm(res) / np.linalg.norm(b) < 1e-10


def test_clover_gamma5_hermiticity_and_gauge_covariance():
    """g5 M(mu) g5 = M(-mu)^dag for the full twisted-clover operator; gauge covariance of the clover term"""
    X = X4
    o, U, Ubc, gq, clov = _clover_setup(X)
    o.set_clover(clov)
    psi = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=5), X)
    g5 = o.gamma5()
    def G5(v):
        return lu.c2r(np.einsum("st,xtc->xsc", g5, lu.r2c(v)))
    a = G5(o.mat(gq, G5(psi), KAPPA, MU, 0))
    b = o.mat(gq, psi, KAPPA, -MU, 1)
    assert lu.rel_l2(a, b) < 1e-14
    o.set_clover(None)
    G = lu.random_su3_lex(X, seed=12)[0]
    xs, ys, zs, ts = lu.coords_lex(X)
    Urot = np.empty_like(Ubc)
    for mu in range(4):
        c = [xs.copy(), ys.copy(), zs.copy(), ts.copy()]
        c[mu] = (c[mu] + 1) % X[mu]
        nb = c[0] + X[0] * (c[1] + X[1] * (c[2] + X[2] * c[3]))
        Urot[mu] = np.einsum("xab,xbc,xdc->xad", G, Ubc[mu], np.conj(G[nb]))
    C0 = lu.dense_clover(Ubc, X, gamma_ukqcd(), 0.2)
    C1 = lu.dense_clover(Urot, X, gamma_ukqcd(), 0.2)
    want = np.einsum("xab,xsbtc,xdc->xsatd", G, C0, np.conj(G))
    assert np.abs(C1 - want).max() < 1e-13
