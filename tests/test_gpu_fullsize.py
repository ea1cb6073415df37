"""GPU parity at BASELINE.json's full single-GPU size, 48^3 x 96 (configs[3]): the device against the CPU oracle on the whole
lattice (the oracle needs about 0.4 s per hop on the box's cores), plus the size-independent properties of the operator --
adjointness, gamma5-hermiticity, linearity, fused-vs-unfused composition, the CG recurrence -- evaluated with device
reductions.  One context for the whole module (gauge upload 6 GB)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X = (48, 48, 48, 96)
KAPPA = 1.0 / (2.0 * (4.0 + 0.1))
MU = 0.1


@pytest.fixture(scope="module")
def S():
    import tmq as T
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    from oracle.oracle import Oracle

    class St:
        pass
    s = St()
    s.T = T
    s.gauge = T.gen_gauge(X, seed=137, t_boundary=-1)
    s.Vh = int(np.prod(X)) // 2
    full = T.gen_spinor(X, "gaussian", seed=101)
    s.even = np.ascontiguousarray(full[: s.Vh]); s.odd = np.ascontiguousarray(full[s.Vh:])
    del full
    s.ctx = T.Context(X)
    s.ctx.load_gauge(s.gauge, t_boundary=-1, recon=12)
    s.ctx.set_op(KAPPA, MU, T.MATPC_EVEN_EVEN)
    s.orc = Oracle(X)
    yield s
    s.ctx.close()


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def test_hop_and_mdagm_match_the_oracle_on_the_whole_lattice(S):
    c, o = S.ctx, S.orc
    a, b = c.spinor(), c.spinor()
    a.set(S.odd)
    c.dslash(b, a, 0, 0)
    assert rel(b.get(), o.dslash(S.gauge, S.odd, 0, 0)) < 1e-13
    a.set(S.even)
    c.dslash(b, a, 1, 1)
    assert rel(b.get(), o.dslash(S.gauge, S.even, 1, 1)) < 1e-13
    c.mdagm(b, a)
    assert rel(b.get(), o.mdagm(S.gauge, S.even, KAPPA, MU, 0)) < 2e-13
    a4, b4 = c.spinor(4), c.spinor(4)
    a4.set(S.even)
    c.mdagm(b4, a4)
    assert rel(b4.get(), o.mdagm(S.gauge, S.even, KAPPA, MU, 0)) < 2e-5


def test_operator_identities_at_full_size(S):
    """<x, M y> = <M^dag x, y>; g5 M_full(mu) g5 = M_full(-mu)^dag; linearity; M^dag M from the fused launches equals the
    composition of the bare hop with separate twists and axpys"""
    c, T = S.ctx, S.T
    x, y, t, u, w = (c.spinor() for _ in range(5))
    x.set(S.even); y.set(S.odd)        # two independent parity-sized fields
    c.matpc(t, y, 0)
    lhs = c.cdot(x, t)
    c.matpc(u, x, 1)
    rhs = c.cdot(u, y)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    # linearity of M^dag M
    c.mdagm(t, x); c.mdagm(u, y)
    c.copy(w, x); c.axpby(-0.7, y, 1.3, w)                       # w = 1.3 x - 0.7 y
    r = c.spinor()
    c.mdagm(r, w)
    c.axpby(1.3, t, 0.0, w); c.axpy(-0.7, u, w)                  # w = 1.3 Mx - 0.7 My
    c.axpy(-1.0, r, w)
    assert c.norm2(w) < (1e-13 ** 2) * c.norm2(r)
    # unfused composition of M = 1 - k^2 A^-1 D A^-1 D from the bare hop: A^-1 = (1 - i a g5)/(1 + a^2)
    a = 2.0 * KAPPA * MU
    def Ainv(dst, src):                                          # dst = A^-1 src, through gamma5 + axpy
        c.copy(dst, src)
        g = c.spinor(); c.copy(g, src); c.gamma5(g)
        c.caxpy(complex(0.0, -a), g, dst)
        c.ax(1.0 / (1.0 + a * a), dst)
        g.free()
    h1, h2, m1 = c.spinor(), c.spinor(), c.spinor()
    c.dslash(h1, x, 1, 0); Ainv(h2, h1)
    c.dslash(h1, h2, 0, 0); Ainv(h2, h1)
    c.copy(m1, x); c.axpy(-KAPPA * KAPPA, h2, m1)                # m1 = M x, unfused
    c.matpc(h1, x, 0)
    c.axpy(-1.0, h1, m1)
    assert c.norm2(m1) < (1e-13 ** 2) * c.norm2(h1)
    # gamma5-hermiticity of the full operator
    fx, fy, fz = c.spinor(8, T.FULL), c.spinor(8, T.FULL), c.spinor(8, T.FULL)
    fx.set(np.concatenate([S.even, S.odd]))
    c.copy(fy, fx); c.gamma5(fy)
    c.mat_full(fz, fy, 0); c.gamma5(fz)                          # g5 M(mu) g5 x
    c.set_op(KAPPA, -MU, T.MATPC_EVEN_EVEN)
    c.mat_full(fy, fx, 1)                                        # M(-mu)^dag x
    c.set_op(KAPPA, MU, T.MATPC_EVEN_EVEN)
    c.axpy(-1.0, fy, fz)
    assert c.norm2(fz) < (1e-13 ** 2) * c.norm2(fy)
    for f in (fx, fy, fz, x, y, t, u, w, r, h1, h2, m1):
        f.free()


def test_cg_history_matches_the_cpu_cg_at_full_size(S):
    """the first iterations of CG on M^dag M reproduce the CPU oracle's residual history (same recurrence, same operator), and
    the complete solve meets the true-residual criterion recomputed through the unfused operator"""
    c, o = S.ctx, S.orc
    b, x = c.spinor(), c.spinor()
    b.set(S.even)
    n = 6
    _, it, _, hist = o.cg_mdagm(S.gauge, S.even, KAPPA, MU, 0, tol=1e-30, maxiter=n)
    c.cg_mdagm(x, b, tol=1e-30, maxiter=n)
    got = c.cg_history(n + 1)
    assert np.allclose(got, hist[: n + 1], rtol=1e-10, atol=0), (got, hist)
    info = c.cg_mdagm(x, b, tol=1e-9, maxiter=2000)
    r = c.spinor()
    c.mdagm(r, x)
    c.axpy(-1.0, b, r)
    assert np.sqrt(c.norm2(r) / c.norm2(b)) <= 1.05e-9 and info["true_res"] <= 1.05e-9 and 10 < info["iter"] < 200
