"""GPU parity tests of the eigensolver layer through the C ABI: the Chebyshev-accelerated
M^dag M (QKXTM_Deflation::polynomialOperator, reference lib/qudaQKXTM_Deflation.cpp:997-1063), the eigensolver
(eigenSolver, :1069-1475) and the exact-deflation projector (deflateVector, :614-800), against the CPU oracle's
restatement of the recurrence and ARPACK (scipy) on the oracle's operator."""
import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

KAPPA = 1.0 / (2.0 * (4.0 + 0.1))
MU = 0.1
X_SMALL = (4, 4, 4, 8)


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return T


def _cplx(a):
    return np.ascontiguousarray(a[..., 0] + 1j * a[..., 1]).ravel()


def _real(v, Vh):
    return np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(Vh, 4, 3, 2))


class Setup:
    def __init__(self, T, X, matpc, part=None):
        from oracle.oracle import Oracle
        self.X, self.matpc = X, matpc
        self.orc = Oracle(X)
        self.Vh = self.orc.Vh
        self.gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=-1)
        self.ctx = T.Context(X)
        if part is not None:
            self.ctx.force_partition(part)
        self.ctx.load_gauge(self.gauge, t_boundary=-1, recon=12)
        self.ctx.set_op(KAPPA, MU, matpc)
        psi = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)
        self.even = np.ascontiguousarray(psi[: self.Vh])

    def A(self, v):
        """oracle M^dag M on a complex vector of length 12 Vh"""
        from oracle.oracle import Oracle
        self.orc = Oracle(self.X)     # the C oracle keeps the lattice in globals
        return _cplx(self.orc.mdagm(self.gauge, _real(v, self.Vh), KAPPA, MU, self.matpc))


@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
@pytest.mark.parametrize("deg", [0, 1, 2, 3, 8, 25])
def test_polynomial_operator_matches_oracle(tmq, matpc, deg):
    from oracle.oracle import poly_operator
    s = Setup(tmq, (4, 6, 4, 8), matpc)
    c = s.ctx
    amin, amax = 0.3, 2.0
    ref = poly_operator(s.A, _cplx(s.even), deg, amin, amax)
    a, b = c.spinor(), c.spinor()
    a.set(s.even)
    c.poly_mdagm(b, a, deg, amin, amax)
    assert lu.rel_l2(_cplx(b.get()), ref) < 5e-13, (matpc, deg)
    # the input survives
    assert np.array_equal(a.get(), s.even)
    c.close()


def test_polynomial_operator_fp32(tmq):
    from oracle.oracle import poly_operator
    s = Setup(tmq, (4, 6, 4, 8), 0)
    c = s.ctx
    ref = poly_operator(s.A, _cplx(s.even), 10, 0.3, 2.0)
    a, b = c.spinor(tmq.PREC_SINGLE), c.spinor(tmq.PREC_SINGLE)
    a.set(s.even)
    c.poly_mdagm(b, a, 10, 0.3, 2.0)
    assert lu.rel_l2(_cplx(b.get()), ref) < 2e-5
    c.close()


# spectrum of M^dag M on this lattice / gauge field: [0.3496, 1.887], eigenvalues 8 and 9 at 0.3713 and 0.3981
AMIN, AMAX = 0.385, 2.0


@pytest.mark.parametrize("matpc,deg,nkv,part", [(2, 30, 32, None), (0, 30, 32, None), (2, 6, 20, None), (2, 0, 40, None),
                                                (2, 6, 20, (0, 0, 1, 1)), (1, 12, 24, (0, 0, 0, 1))])
def test_eigensolver_matches_arpack_on_the_oracle(tmq, matpc, deg, nkv, part):
    """smallest eigenpairs of M^dag M: eigenvalues vs ARPACK on the oracle operator (Chebyshev-accelerated exactly as the
    reference's isACC branch), residuals with the oracle operator, orthonormality, subspace overlap"""
    from oracle.oracle import eigs_reference
    s = Setup(tmq, X_SMALL, matpc, part)
    c = s.ctx
    nev = 8
    amin, amax = AMIN, AMAX
    n = 12 * s.Vh
    poly = (deg, amin, amax) if deg > 0 else None
    lam_ref, U_ref = eigs_reference(s.A, n, nev, nkv, "SR", poly=poly, tol=1e-12)
    es = c.eigset(nkv + 1)
    r = c.eigensolve(es, nev, nkv, poly_deg=deg, amin=amin, amax=amax, tol=1e-12, max_restarts=400, which=0, seed=7)
    assert r["nconv"] == nev, r
    assert np.allclose(r["evals"], lam_ref, rtol=1e-10, atol=0), (r["evals"], lam_ref)
    V = np.stack([_cplx(es.vector(i).get()) for i in range(nev)], axis=1)
    G = V.conj().T @ V
    assert np.abs(G - np.eye(nev)).max() < 1e-11
    for i in range(nev):
        res = np.linalg.norm(s.A(V[:, i]) - r["evals"][i] * V[:, i])
        assert res < 1e-9 and abs(res - r["resid"][i]) < 1e-9, (i, res, r["resid"][i])
    # same invariant subspace as ARPACK's (degenerate pairs may be rotated inside it)
    sv = np.linalg.svd(U_ref.conj().T @ V, compute_uv=False)
    assert sv.min() > 1 - 1e-8, sv
    c.close()


def test_eigensolver_largest(tmq):
    from oracle.oracle import eigs_reference
    s = Setup(tmq, X_SMALL, 2)
    c = s.ctx
    nev, nkv = 4, 24
    lam_ref, _ = eigs_reference(s.A, 12 * s.Vh, nev, nkv, "LR", tol=1e-12)
    es = c.eigset(nkv + 1)
    r = c.eigensolve(es, nev, nkv, tol=1e-12, max_restarts=400, which=1, seed=3)
    assert r["nconv"] == nev and np.allclose(r["evals"], lam_ref, rtol=1e-10)
    c.close()


def test_deflation_projector(tmq):
    """out = U Lambda^-1 U^dag in (deflateVector) vs numpy on the device's own eigenvectors; and, as an initial guess,
    it removes the low-mode part of the CG residual: |b - A x0| has no component along the eigenvectors"""
    s = Setup(tmq, X_SMALL, 2)
    c = s.ctx
    nev, nkv = 6, 28
    es = c.eigset(nkv + 1)
    r = c.eigensolve(es, nev, nkv, poly_deg=20, amin=AMIN, amax=AMAX, tol=1e-12, max_restarts=400, which=0, seed=11)
    V = np.stack([_cplx(es.vector(i).get()) for i in range(nev)], axis=1)
    b = _cplx(s.even)
    ref = V @ ((V.conj().T @ b) / r["evals"])
    a, x0 = c.spinor(), c.spinor()
    a.set(s.even)
    c.deflate(x0, a, es, r["evals"])
    got = _cplx(x0.get())
    assert lu.rel_l2(got, ref) < 1e-12
    resid = b - s.A(got)
    assert np.abs(V.conj().T @ resid).max() < 1e-8 * np.linalg.norm(b)
    # projectVector: out = in - U U^dag in (here on the first 4 vectors), in place as well
    pr = c.spinor()
    c.project(pr, a, es, 4)
    ref_p = b - V[:, :4] @ (V[:, :4].conj().T @ b)
    assert lu.rel_l2(_cplx(pr.get()), ref_p) < 1e-13
    c.project(a, a, es, 4)
    assert lu.rel_l2(_cplx(a.get()), ref_p) < 1e-13
    c.close()


# ---- the unpreconditioned operator (isFullOp: what calc_loops deflates with, lib/qudaQKXTM_interface.cpp:1725-1736) ----
class FullSetup(Setup):
    def __init__(self, T, X, part=None):
        super().__init__(T, X, 0, part)
        self.full = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)

    def A(self, v):
        """oracle M_full^dag M_full on a complex vector of length 24 Vh ([even | odd] order)"""
        from oracle.oracle import Oracle
        self.orc = Oracle(self.X)
        x = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(2 * self.Vh, 4, 3, 2))
        y = self.orc.mat(self.gauge, x, KAPPA, MU, 0)
        return _cplx(self.orc.mat(self.gauge, y, KAPPA, MU, 1))


@pytest.mark.parametrize("deg", [-1, 1, 2, 7])
def test_full_operator_polynomial(tmq, deg):
    from oracle.oracle import poly_operator
    s = FullSetup(tmq, (4, 6, 4, 8))
    c = s.ctx
    a, b = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL)
    a.set(s.full)
    c.poly_mdagm(b, a, deg, 0.05, 3.5)
    ref = s.A(_cplx(s.full)) if deg < 0 else poly_operator(s.A, _cplx(s.full), deg, 0.05, 3.5)
    assert lu.rel_l2(_cplx(b.get()), ref) < 5e-13, deg
    c.close()


@pytest.mark.parametrize("part", [None, (0, 0, 0, 1)])
def test_full_operator_eigensolver_and_projection(tmq, part):
    """isFullOp: smallest eigenpairs of M_full^dag M_full vs ARPACK on the oracle's full operator, and projectVector
    (x - U U^dag x, lib/qudaQKXTM_Deflation.cpp:1926-2060) on FULL fields"""
    from oracle.oracle import eigs_reference
    s = FullSetup(tmq, X_SMALL, part)
    c = s.ctx
    nev, nkv = 6, 30
    n = 24 * s.Vh
    lam_ref, U_ref = eigs_reference(s.A, n, nev, nkv, "SR", tol=1e-12)
    # Chebyshev window from the reference spectrum: just above the wanted eigenvalues, up to beyond the largest one
    amin, amax = 1.6 * lam_ref[-1], 3.5
    es = c.eigset(nkv + 1, tmq.PREC_DOUBLE, tmq.FULL)
    r = c.eigensolve(es, nev, nkv, poly_deg=24, amin=amin, amax=amax, tol=1e-12, max_restarts=400, which=0, seed=7)
    assert r["nconv"] == nev, r
    assert np.allclose(r["evals"], lam_ref, rtol=1e-9, atol=0), (r["evals"], lam_ref)
    V = np.stack([_cplx(es.vector(i).get()) for i in range(nev)], axis=1)
    assert np.abs(V.conj().T @ V - np.eye(nev)).max() < 1e-11
    sv = np.linalg.svd(U_ref.conj().T @ V, compute_uv=False)
    assert sv.min() > 1 - 1e-8, sv
    x, y = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL)
    x.set(s.full)
    c.project(y, x, es, nev)
    b = _cplx(s.full)
    assert lu.rel_l2(_cplx(y.get()), b - V @ (V.conj().T @ b)) < 1e-13
    c.deflate(y, x, es, r["evals"])
    assert lu.rel_l2(_cplx(y.get()), V @ ((V.conj().T @ b) / r["evals"])) < 1e-12
    c.close()
