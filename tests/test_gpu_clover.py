"""GPU parity tests of the twisted-clover variant through the C ABI against the CPU oracle with its
dense 12x12 clover matrices: the device builds C and (C + i a g5)^-1 itself from the resident gauge field
(tmq_clover_load = loadCloverQuda(NULL, NULL, &inv_param), qkxtm/MG_Bench.cpp:605-608), so these tests cover the clover
leaves, the chiral-basis storage, the 6x6 block inverse and every fused epilogue.  Tolerances as for twisted mass."""
import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

KAPPA = 1.0 / (2.0 * (4.0 + 0.1))
MU = 0.1
CSW = 1.57551
X = (4, 6, 4, 8)
TOL = {8: 2e-13, 4: 2e-5}


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return T


class Setup:
    def __init__(self, T, recon, csw=CSW, mu=MU):
        from oracle.oracle import Oracle
        self.orc = Oracle(X)
        self.Vh = self.orc.Vh
        self.gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=-1)
        self.clov = self.orc.clover_compute(self.gauge, csw * KAPPA)
        self.ctx = T.Context(X)
        self.ctx.load_gauge(self.gauge, t_boundary=-1, recon=recon)
        self.ctx.set_op(KAPPA, mu, 0)
        self.ctx.clover_load(csw * KAPPA)
        self.full = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)
        self.even = np.ascontiguousarray(self.full[: self.Vh]); self.odd = np.ascontiguousarray(self.full[self.Vh:])

    def oracle(self):
        from oracle.oracle import Oracle
        self.orc = Oracle(X)
        self.orc.set_clover(self.clov)
        return self.orc


@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
def test_full_operator_and_site_terms(tmq, recon, prec):
    s = Setup(tmq, recon); o = s.oracle(); c = s.ctx
    for dagger in (0, 1):
        a, b = c.spinor(prec, tmq.FULL), c.spinor(prec, tmq.FULL)
        a.set(s.full)
        c.mat_full(b, a, dagger)
        assert lu.rel_l2(b.get(), o.mat(s.gauge, s.full, KAPPA, MU, dagger)) < TOL[prec], dagger
    # A^-1 D (dagger 0) and A^-dag D^dag (dagger 1) on both parities, with and without the xpay term
    a, b, x = c.spinor(prec), c.spinor(prec), c.spinor(prec)
    for out_parity in (0, 1):
        src = s.odd if out_parity == 0 else s.even
        xin = s.even if out_parity == 0 else s.odd
        a.set(src); x.set(xin)
        for dagger in (0, 1):
            hop = o.dslash(s.gauge, src, out_parity, dagger)
            ref = o.site_A(hop, KAPPA, MU, out_parity, dagger, 1)
            c.dslash_twist_xpay(b, a, out_parity, dagger)
            assert lu.rel_l2(b.get(), ref) < TOL[prec], (out_parity, dagger)
            c.dslash_twist_xpay(b, a, out_parity, dagger, x=x, k=-0.37)
            assert lu.rel_l2(b.get(), xin - 0.37 * ref) < TOL[prec], (out_parity, dagger)
    c.close()


@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
@pytest.mark.parametrize("prec", [8, 4])
def test_matpc_mdagm_poly(tmq, matpc, prec):
    from oracle.oracle import poly_operator
    s = Setup(tmq, 12); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, matpc)
    a, b = c.spinor(prec), c.spinor(prec)
    a.set(s.even)
    for dagger in (0, 1):
        c.matpc(b, a, dagger)
        assert lu.rel_l2(b.get(), o.matpc(s.gauge, s.even, KAPPA, MU, matpc, dagger)) < TOL[prec], dagger
    c.mdagm(b, a)
    assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, MU, matpc)) < 2 * TOL[prec]
    if prec == 8:
        cplx = lambda v: np.ascontiguousarray(v[..., 0] + 1j * v[..., 1]).ravel()
        real = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(s.Vh, 4, 3, 2))
        A = lambda v: cplx(o.mdagm(s.gauge, real(v), KAPPA, MU, matpc))
        c.poly_mdagm(b, a, 6, 0.3, 2.5)
        assert lu.rel_l2(cplx(b.get()), poly_operator(A, cplx(s.even), 6, 0.3, 2.5)) < 1e-12
    c.close()


@pytest.mark.parametrize("matpc", [0, 3])
def test_solve_full_system(tmq, matpc):
    """prepare -> M^dag -> CG on M^dag M -> reconstruct (lib/qudaQKXTM_interface.cpp:2020-2041) with the twisted-clover
    operator: iteration count vs the CPU CG, true residual, and the full-lattice residual with the oracle operator"""
    s = Setup(tmq, 12); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, matpc)
    bsrc = lu.spinor_eo_from_lex(lu.z4_source_lex(X, seed=100), X)
    b, x = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL)
    spc, rhs, xpc = c.spinor(), c.spinor(), c.spinor()
    b.set(bsrc)
    c.prepare(spc, b)
    assert lu.rel_l2(spc.get(), o.prepare(s.gauge, bsrc, KAPPA, MU, matpc)) < 1e-13
    c.matpc(rhs, spc, 1)
    info = c.cg_mdagm(xpc, rhs, tol=1e-10, maxiter=3000)
    _, it_ref, _, _ = o.cg_mdagm(s.gauge, rhs.get(), KAPPA, MU, matpc, tol=1e-10, maxiter=3000)
    assert abs(info["iter"] - it_ref) <= 2 and info["true_res"] <= 1.05e-10, (info, it_ref)
    c.reconstruct(x, xpc, b)
    res = o.mat(s.gauge, x.get(), KAPPA, MU, 0) - bsrc
    assert np.linalg.norm(res) / np.linalg.norm(bsrc) < 1e-8
    if matpc == 0:
        info4 = c.cg_mdagm(xpc, rhs, tol=1e-10, maxiter=3000, sloppy_prec=4, reliable_delta=0.1)
        assert info4["true_res"] <= 1.05e-10
    c.close()


def test_csw_zero_is_twisted_mass_and_mu_change_rebuilds_inverse(tmq):
    s = Setup(tmq, 12, csw=0.0); c = s.ctx
    from oracle.oracle import Oracle
    o = Oracle(X)                                   # no clover term
    a, b = c.spinor(), c.spinor()
    a.set(s.even)
    c.mdagm(b, a)
    assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, MU, 0)) < 2e-13
    c.close()
    s = Setup(tmq, 12); o = s.oracle(); c = s.ctx
    a, b = c.spinor(), c.spinor()
    a.set(s.even)
    for mu in (MU, -MU, 0.03):
        c.set_op(KAPPA, mu, 0)
        c.mdagm(b, a)
        assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, mu, 0)) < 4e-13, mu
    c.clover_free()
    c.mdagm(b, a)
    o.set_clover(None)
    assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, 0.03, 0)) < 2e-13
    c.close()


@pytest.mark.parametrize("part", [(0, 0, 0, 1), (0, 0, 1, 0), (0, 0, 1, 1)])
def test_clover_on_the_ghost_zone_path(tmq, part):
    """tmq_force_partition routes z / t through the ghost-zone machinery on one GPU (the reference's --partition): the clover
    term is then built from the gauge field with an exchanged one-site halo (incl. the z-t corners) and the fused sharded
    Dslash launches run their clover variants; results must equal the oracle on the periodic lattice"""
    from oracle.oracle import Oracle
    orc = Oracle(X)
    gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=-1)
    clov = orc.clover_compute(gauge, CSW * KAPPA)
    full = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)
    even = np.ascontiguousarray(full[: orc.Vh])
    for p2p in (4, 2, 0):
        c = tmq.Context(X)
        c.force_partition(part)
        c.set_option(tmq.OPT_HALO_P2P, p2p)
        c.load_gauge(gauge, t_boundary=-1, recon=12)
        c.set_op(KAPPA, MU, 0)
        c.clover_load(CSW * KAPPA)
        orc = Oracle(X); orc.set_clover(clov)
        a, b = c.spinor(), c.spinor()
        a.set(even)
        c.mdagm(b, a)
        assert lu.rel_l2(b.get(), orc.mdagm(gauge, even, KAPPA, MU, 0)) < 4e-13, (part, p2p)
        fa, fb = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL)
        fa.set(full)
        c.mat_full(fb, fa, 0)
        assert lu.rel_l2(fb.get(), orc.mat(gauge, full, KAPPA, MU, 0)) < 2e-13, (part, p2p)
        x = c.spinor()
        info = c.cg_mdagm(x, a, tol=1e-9, maxiter=2000)
        _, it_ref, _, _ = orc.cg_mdagm(gauge, even, KAPPA, MU, 0, tol=1e-9, maxiter=2000)
        assert abs(info["iter"] - it_ref) <= 2 and info["true_res"] <= 1.05e-9
        c.close()
