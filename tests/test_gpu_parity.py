"""GPU parity tests: the sm_100a path, called through the C ABI (include/tmq.h, via ctypes), against the CPU
oracle on identical seeded inputs.  Modelled on upstream's dslash_test / invert_test / blas_test (SURVEY.md
section 4): random SU(3) gauge field, random spinor, one application on the device, one on the host,
relative L2 difference.  Tolerances are BASELINE.json's: single application 1e-13 (fp64) / 1e-5 (fp32);
CG iteration count within +-2 of the CPU CG, true residual <= tol."""
import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

KAPPA = 1.0 / (2.0 * (4.0 + 0.1))   # default mass 0.1 (qkxtm/Calc_Loops.cpp:382-388)
MU = 0.1
TOL = {8: 1e-13, 4: 1e-5}
LATTICES = [(8, 8, 8, 16), (4, 6, 4, 8)]


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return T


class Setup:
    def __init__(self, T, X, recon, t_boundary=-1):
        from oracle.oracle import Oracle
        self.X = X
        self.orc = Oracle(X)
        self.gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=t_boundary)
        self.ctx = T.Context(X)
        self.ctx.load_gauge(self.gauge, t_boundary=t_boundary, recon=recon)
        psi = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)
        self.full = psi
        self.Vh = self.orc.Vh
        self.even, self.odd = np.ascontiguousarray(psi[: self.Vh]), np.ascontiguousarray(psi[self.Vh:])

    def oracle(self):
        # the C oracle keeps the lattice in globals; re-arm it for this lattice
        from oracle.oracle import Oracle
        self.orc = Oracle(self.X)
        return self.orc


_cache = {}


def get_setup(T, X, recon):
    key = (X, recon)
    if key not in _cache:
        _cache[key] = Setup(T, X, recon)
    return _cache[key]


@pytest.mark.parametrize("X", LATTICES)
@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
def test_hop_matches_oracle(tmq, X, recon, prec):
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    for out_parity in (0, 1):
        src = s.odd if out_parity == 0 else s.even
        for dagger in (0, 1):
            ref = o.dslash(s.gauge, src, out_parity, dagger)
            a, b = c.spinor(prec), c.spinor(prec)
            a.set(src)
            c.dslash(b, a, out_parity, dagger)
            assert lu.rel_l2(b.get(), ref) < TOL[prec], (out_parity, dagger)


@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("mu", [MU, -MU])
def test_tm_dslash_and_xpay_match_oracle(tmq, recon, prec, mu):
    X = LATTICES[0]
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, mu, tmq.MATPC_EVEN_EVEN)
    a, b, x = c.spinor(prec), c.spinor(prec), c.spinor(prec)
    a.set(s.even); x.set(s.odd)
    # K2: out = A^-1 D in ; dagger: A^-dag D^dag in
    ref = o.tm_dslash(s.gauge, s.even, KAPPA, mu, 1, 0)
    c.dslash_twist_xpay(b, a, 1, 0)
    assert lu.rel_l2(b.get(), ref) < TOL[prec]
    ref_d = o.twist(o.dslash(s.gauge, s.even, 1, 1), KAPPA, mu, dagger=1, inverse=1)
    c.dslash_twist_xpay(b, a, 1, 1)
    assert lu.rel_l2(b.get(), ref_d) < TOL[prec]
    # K3: out = x + k A^-1 D in
    k = -KAPPA * KAPPA
    c.dslash_twist_xpay(b, a, 1, 0, x=x, k=k)
    assert lu.rel_l2(b.get(), s.odd + k * ref) < TOL[prec]


@pytest.mark.parametrize("X", LATTICES)
@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("matpc", [0, 1, 2, 3])
def test_matpc_and_mdagm_match_oracle(tmq, X, recon, prec, matpc):
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, matpc)
    src = s.even if (matpc & 1) == 0 else s.odd
    a, b = c.spinor(prec), c.spinor(prec)
    a.set(src)
    for dagger in (0, 1):
        ref = o.matpc(s.gauge, src, KAPPA, MU, matpc, dagger)
        c.matpc(b, a, dagger)
        assert lu.rel_l2(b.get(), ref) < TOL[prec], dagger
    ref = o.mdagm(s.gauge, src, KAPPA, MU, matpc)
    c.mdagm(b, a)
    assert lu.rel_l2(b.get(), ref) < 2 * TOL[prec]


@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
def test_full_operator_prepare_reconstruct(tmq, recon, prec):
    X = LATTICES[1]
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    for matpc in (0, 1, 2):
        c.set_op(KAPPA, -MU, matpc)
        f_in, f_out = c.spinor(prec, tmq.FULL), c.spinor(prec, tmq.FULL)
        f_in.set(s.full)
        for dagger in (0, 1):
            c.mat_full(f_out, f_in, dagger)
            assert lu.rel_l2(f_out.get(), o.mat(s.gauge, s.full, KAPPA, -MU, dagger)) < TOL[prec]
        src = c.spinor(prec)
        c.prepare(src, f_in)
        assert lu.rel_l2(src.get(), o.prepare(s.gauge, s.full, KAPPA, -MU, matpc)) < TOL[prec]
        # reconstruct from an arbitrary parity "solution"
        p = matpc & 1
        xp = s.even if p == 0 else s.odd
        xpc = c.spinor(prec); xpc.set(xp)
        c.reconstruct(f_out, xpc, f_in)
        xfull = np.zeros_like(s.full)
        xfull[p * s.Vh:(p + 1) * s.Vh] = xp
        ref = o.reconstruct(s.gauge, xfull, s.full, KAPPA, -MU, matpc)
        assert lu.rel_l2(f_out.get(), ref) < TOL[prec]


def test_schur_identity_solves_full_system(tmq):
    """prepare -> CG on M^dag M (rhs M^dag src) -> reconstruct reproduces M_full x = b
    (the calc_loops CG branch, lib/qudaQKXTM_interface.cpp:2020-2041)"""
    X = LATTICES[1]
    s = get_setup(tmq, X, 12); c = s.ctx
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    b = c.spinor(8, tmq.FULL); b.set(s.full)
    src, rhs, xpc = c.spinor(), c.spinor(), c.spinor()
    c.prepare(src, b)
    c.matpc(rhs, src, 1)                       # in <- M^dag in (interface.cpp:2034)
    info = c.cg_mdagm(xpc, rhs, tol=1e-12, maxiter=2000)
    assert info["true_res"] < 1e-11
    x = c.spinor(8, tmq.FULL)
    c.reconstruct(x, xpc, b)
    chk = c.spinor(8, tmq.FULL)
    c.mat_full(chk, x, 0)
    assert lu.rel_l2(chk.get(), s.full) < 1e-10


@pytest.mark.parametrize("X,tol", [(LATTICES[0], 1e-7), (LATTICES[0], 1e-10), (LATTICES[1], 1e-9)])
@pytest.mark.parametrize("recon", [12, 18])
def test_cg_iterations_and_residual_match_cpu_cg(tmq, X, tol, recon):
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    z4 = lu.spinor_eo_from_lex(lu.z4_source_lex(X, seed=100), X)
    rhs = np.ascontiguousarray(z4[: s.Vh])
    x_ref, it_ref, tr_ref, hist_ref = o.cg_mdagm(s.gauge, rhs, KAPPA, MU, 0, tol=tol, maxiter=5000)
    b, x = c.spinor(), c.spinor()
    b.set(rhs)
    info = c.cg_mdagm(x, b, tol=tol, maxiter=5000)
    assert abs(info["iter"] - it_ref) <= 2, (info, it_ref)
    assert info["true_res"] <= tol * 1.05
    assert abs(info["true_res"] - tr_ref) <= 0.1 * tr_ref + 1e-15
    assert lu.rel_l2(x.get(), x_ref) < 10 * tol
    hist = c.cg_history(info["iter"] + 1)
    n = min(len(hist), len(hist_ref), 20)
    assert np.allclose(hist[:n], hist_ref[:n], rtol=1e-9)


def test_cg_asymmetric_generic_path(tmq):
    X = LATTICES[1]
    s = get_setup(tmq, X, 18); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, tmq.MATPC_ODD_ODD_ASYM)
    x_ref, it_ref, tr_ref, _ = o.cg_mdagm(s.gauge, s.odd, KAPPA, MU, 3, tol=1e-9, maxiter=5000)
    b, x = c.spinor(), c.spinor()
    b.set(s.odd)
    info = c.cg_mdagm(x, b, tol=1e-9, maxiter=5000)
    assert abs(info["iter"] - it_ref) <= 2
    assert info["true_res"] <= 1.05e-9
    assert lu.rel_l2(x.get(), x_ref) < 1e-8


@pytest.mark.parametrize("recon", [12, 18])
def test_mixed_precision_cg_reaches_fp64_residual(tmq, recon):
    X = LATTICES[0]
    s = get_setup(tmq, X, recon); o = s.oracle(); c = s.ctx
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    x_ref, it_ref, _, _ = o.cg_mdagm(s.gauge, s.even, KAPPA, MU, 0, tol=1e-10, maxiter=5000)
    b, x = c.spinor(), c.spinor()
    b.set(s.even)
    # the reference drivers' setting (qkxtm/Calc_Loops.cpp:481): the same iteration count as the fp64 CPU CG, +-2
    info = c.cg_mdagm(x, b, tol=1e-10, maxiter=5000, reliable_delta=1e-4, sloppy_prec=4)
    assert info["true_res"] <= 1.05e-10
    assert abs(info["iter"] - it_ref) <= 2, (info["iter"], it_ref)
    assert lu.rel_l2(x.get(), x_ref) < 1e-9
    # a loose delta (an fp64 residual recomputation every decade) still converges, at a bounded cost in iterations
    info = c.cg_mdagm(x, b, tol=1e-10, maxiter=5000, reliable_delta=1e-1, sloppy_prec=4)
    assert info["true_res"] <= 1.05e-10
    assert info["iter"] <= int(1.5 * it_ref) + 10
    assert lu.rel_l2(x.get(), x_ref) < 1e-8


@pytest.mark.parametrize("part", [(0, 0, 0, 1), (0, 0, 1, 0), (0, 0, 1, 1)])
@pytest.mark.parametrize("recon", [12, 18])
@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("p2p", [4, 3, 2, 1, 0])
def test_ghost_zone_path_equals_periodic_path(tmq, part, recon, prec, p2p):
    """--partition style self-exchange (qkxtm/QKXTM_util.cpp:1717-1720): pack -> exchange -> interior +
    boundary launches must reproduce the single-launch result and the oracle."""
    X = LATTICES[1]
    s = get_setup(tmq, X, recon); o = s.oracle()
    c = tmq.Context(X)
    c.force_partition(part)
    # p2p = 1 / 2: peer-memory halo (faces stored by the pack kernel / pushed by the copy engines into the neighbour's
    # ghost arena, one fused launch whose boundary CTAs wait on arrival flags); 0: pack -> exchange stream ->
    # interior + boundary launches
    c.set_option(tmq.OPT_HALO_P2P, p2p)
    assert c.halo_mode() == 1 + p2p
    c.load_gauge(s.gauge, t_boundary=-1, recon=recon)
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    a, b = c.spinor(prec), c.spinor(prec)
    for out_parity in (0, 1):
        src = s.odd if out_parity == 0 else s.even
        a.set(src)
        for dagger in (0, 1):
            c.dslash(b, a, out_parity, dagger)
            assert lu.rel_l2(b.get(), o.dslash(s.gauge, src, out_parity, dagger)) < TOL[prec]
    a.set(s.even)
    c.mdagm(b, a)
    assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, MU, 0)) < 2 * TOL[prec]
    # p2p = 3: fused compute + halo exchange -- inside M_pc / M^dag M / the CG iteration the boundary CTAs of the launch that produces
    # a field send its faces for the next application themselves (both matpc directions and daggers, the asymmetric operator too)
    for dagger in (0, 1):
        c.matpc(b, a, dagger)
        assert lu.rel_l2(b.get(), o.matpc(s.gauge, s.even, KAPPA, MU, 0, dagger)) < 2 * TOL[prec]
    c.set_op(KAPPA, MU, tmq.MATPC_ODD_ODD_ASYM)
    a.set(s.odd)
    for dagger in (0, 1):
        c.matpc(b, a, dagger)
        assert lu.rel_l2(b.get(), o.matpc(s.gauge, s.odd, KAPPA, MU, 3, dagger)) < 2 * TOL[prec]
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    a.set(s.even)
    if prec == 8:
        x_ref, it_ref, _, _ = o.cg_mdagm(s.gauge, s.even, KAPPA, MU, 0, tol=1e-9, maxiter=5000)
        x = c.spinor()
        info = c.cg_mdagm(x, a, tol=1e-9, maxiter=5000)
        assert abs(info["iter"] - it_ref) <= 2 and info["true_res"] <= 1.05e-9
        assert lu.rel_l2(x.get(), x_ref) < 1e-8
        infom = c.cg_mdagm(x, a, tol=1e-9, maxiter=5000, sloppy_prec=4, reliable_delta=1e-4)
        assert abs(infom["iter"] - it_ref) <= 2 and infom["true_res"] <= 1.05e-9
        # faces sent ahead by one loop must never be taken for those of another: the timing loop of bench.py ends with the faces of ITS
        # last search direction in flight, and the solve after it rewrites the same work field (this once gave a converged-looking
        # wrong solution on 8 GPUs); likewise a second solve with another right-hand side
        c.time_kernel(4, prec, 3, a)
        info = c.cg_mdagm(x, a, tol=1e-9, maxiter=5000)
        assert abs(info["iter"] - it_ref) <= 2 and info["true_res"] <= 1.05e-9 and lu.rel_l2(x.get(), x_ref) < 1e-8
        b.set(s.even[::-1].copy())
        c.cg_mdagm(x, b, tol=1e-9, maxiter=5000)
        info = c.cg_mdagm(x, a, tol=1e-9, maxiter=5000)
        assert abs(info["iter"] - it_ref) <= 2 and lu.rel_l2(x.get(), x_ref) < 1e-8
    c.close()


@pytest.mark.parametrize("part", [None, (0, 0, 0, 1), (0, 0, 1, 1)])
def test_cg_host_loop_ahead_of_the_residual_gives_the_same_iterates(tmq, part):
    """TMQ_OPT_CG_LAG: the host enqueues L iterations ahead of the |r|^2 it reads and the device takes the stopping test itself; launches
    enqueued past convergence exit at once.  Iteration count, residual history and solution must be those of the synchronous loop
    (L = 0) bit for bit, also when maxiter cuts the solve short inside the look-ahead window."""
    X = LATTICES[1]
    s = get_setup(tmq, X, 12)
    ref = None
    for lag in (0, 1, 2, 3, 6):
        c = tmq.Context(X)
        if part is not None:
            c.force_partition(part)
        c.set_option(tmq.OPT_CG_LAG, lag)
        c.load_gauge(s.gauge, t_boundary=-1, recon=12)
        c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
        b, x = c.spinor(), c.spinor()
        b.set(s.even)
        info = c.cg_mdagm(x, b, tol=1e-9, maxiter=5000)
        hist = c.cg_history(info["iter"] + 1)                          # |b|^2, then |r|^2 after every iteration
        x1 = x.get().copy()
        cut = c.cg_mdagm(x, b, tol=1e-9, maxiter=info["iter"] - 2)      # stops on maxiter with unread residuals in flight
        x2 = x.get().copy()
        again = c.cg_mdagm(x, b, tol=1e-9, maxiter=5000)                # and the context is fit for the next solve
        got = (info["iter"], info["true_res"], cut["iter"], cut["true_res"], again["iter"])
        if ref is None:
            ref = (got, x1, x2, hist)
        else:
            assert got == ref[0], (lag, got, ref[0])
            assert np.array_equal(x1, ref[1]) and np.array_equal(x2, ref[2]) and np.array_equal(x.get(), ref[1]), lag
            assert np.array_equal(hist, ref[3]), lag
        assert cut["iter"] == info["iter"] - 2
        c.close()


def test_blas_against_numpy(tmq):
    X = LATTICES[1]
    s = get_setup(tmq, X, 18); c = s.ctx
    rng = np.random.default_rng(3)
    for prec, tol in ((8, 1e-14), (4, 1e-6)):
        xs, ys, zs = (rng.standard_normal(s.even.shape) for _ in range(3))
        x, y, z = c.spinor(prec), c.spinor(prec), c.spinor(prec)
        cx = lambda a: lu.r2c(a)
        x.set(xs); y.set(ys); z.set(zs)
        assert abs(c.norm2(x) - np.sum(xs * xs)) < tol * np.sum(xs * xs)
        assert abs(c.redot(x, y) - np.sum(xs * ys)) < tol * np.sum(np.abs(xs * ys))
        ref = np.sum(np.conj(cx(xs)) * cx(ys))
        assert abs(c.cdot(x, y) - ref) < tol * np.sum(np.abs(xs * ys))
        c.axpy(0.3, x, y); ys = ys + 0.3 * xs
        assert lu.rel_l2(y.get(), ys) < tol
        c.axpby(1.5, x, -0.25, y); ys = 1.5 * xs - 0.25 * ys
        assert lu.rel_l2(y.get(), ys) < tol
        c.xpay(x, 0.7, y); ys = xs + 0.7 * ys
        assert lu.rel_l2(y.get(), ys) < tol
        c.ax(-1.25, x); xs = -1.25 * xs
        assert lu.rel_l2(x.get(), xs) < tol
        a = 0.3 - 0.8j
        c.caxpy(a, x, y); ys = lu.c2r(cx(ys) + a * cx(xs))
        assert lu.rel_l2(y.get(), ys) < tol
        b = -0.1 + 0.2j
        c.cxpaypbz(x, a, y, b, z); zs = lu.c2r(cx(xs) + a * cx(ys) + b * cx(zs))
        assert lu.rel_l2(z.get(), zs) < tol
        n = c.axpy_norm(0.5, x, y); ys = ys + 0.5 * xs
        assert abs(n - np.sum(ys * ys)) < 10 * tol * n and lu.rel_l2(y.get(), ys) < tol
        n = c.xmy_norm(x, y); ys = xs - ys
        assert abs(n - np.sum(ys * ys)) < 10 * tol * n and lu.rel_l2(y.get(), ys) < tol
        c.axpy_zpbx(0.2, x, y, z, -0.6); ys = ys + 0.2 * xs; xs = zs - 0.6 * xs
        assert lu.rel_l2(y.get(), ys) < tol and lu.rel_l2(x.get(), xs) < tol
        c.gamma5(x); xs = xs[:, [2, 3, 0, 1]]     # UKQCD gamma5 = spin swap (apply_gamma5_vector_core.h:1-16)
        assert lu.rel_l2(x.get(), xs) < tol
        c.zero(x)
        assert c.norm2(x) == 0.0
    # precision conversion
    d, f = c.spinor(8), c.spinor(4)
    d.set(s.even); c.copy(f, d)
    assert lu.rel_l2(f.get(), s.even) < 1e-7
    c.copy(d, f)
    assert lu.rel_l2(d.get(), s.even) < 1e-7


def test_plaquette_matches_oracle_and_unit_gauge(tmq):
    X = LATTICES[1]
    for recon in (12, 18):
        s = get_setup(tmq, X, recon); o = s.oracle()
        assert abs(s.ctx.plaquette() - o.plaquette(s.gauge)) < 1e-13
    c = tmq.Context(X)
    c.load_gauge(lu.unit_gauge_qdp(X, t_boundary=+1), t_boundary=+1, recon=12)
    assert abs(c.plaquette() - 1.0) < 1e-14
    c.close()


def _qkxtm_from_lex(psi_lex):
    """host AoS [x][s][c][ri] -> QKXTM SoA [(s*3+c)][x][ri] (packVector, lib/qudaQKXTM_Vector.cpp:72-81)"""
    V = psi_lex.shape[0]
    return np.ascontiguousarray(psi_lex.reshape(V, 12, 2).transpose(1, 0, 2))


@pytest.mark.parametrize("qprec", [8, 4])
def test_qkxtm_upload_download_and_container_kernels(tmq, qprec):
    X = LATTICES[1]
    s = get_setup(tmq, X, 12); c = s.ctx
    V = s.orc.V
    qdt = np.float64 if qprec == 8 else np.float32
    lex = lu.gaussian_spinor_lex(X, seed=55)
    qk = _qkxtm_from_lex(lex).astype(qdt)
    dq = c.dev_malloc(qk.nbytes)
    c.h2d(dq, qk)
    eo = lu.spinor_eo_from_lex(lex, X)
    # uploadToCuda: full field, and each parity alone
    f = c.spinor(8, tmq.FULL)
    c.from_qkxtm(f, dq, qprec)
    assert lu.rel_l2(f.get(), eo.astype(qdt)) < 1e-15
    for parity in (0, 1):
        p = c.spinor(8)
        c.from_qkxtm(p, dq, qprec, parity)
        assert lu.rel_l2(p.get(), eo[parity * s.Vh:(parity + 1) * s.Vh].astype(qdt)) < 1e-15
        # downloadFromCuda zero-fills the absent parity and applies the fused 2*kappa rescale
        dq2 = c.dev_malloc(qk.nbytes)
        c.to_qkxtm(dq2, p, qprec, parity, scale=2 * KAPPA)
        back = np.empty_like(qk); c.d2h(back, dq2)
        keep = eo.astype(qdt).astype(np.float64).copy(); keep[(1 - parity) * s.Vh:(2 - parity) * s.Vh] = 0
        ref = _qkxtm_from_lex(lu.spinor_lex_from_eo(keep, X)) * (2 * KAPPA)
        assert lu.rel_l2(back, ref) < (1e-15 if qprec == 8 else 1e-7)
        c.dev_free(dq2)
    # round trip of the full field
    dq3 = c.dev_malloc(qk.nbytes)
    c.to_qkxtm(dq3, f, qprec)
    back = np.empty_like(qk); c.d2h(back, dq3)
    assert np.array_equal(back, qk)
    # scaleVector / apply_gamma5 / casts on the QKXTM layout
    c.qkxtm_scale(dq3, qprec, 3.0); c.d2h(back, dq3)
    assert lu.rel_l2(back, 3.0 * qk) < 1e-7
    c.qkxtm_gamma5(dq3, qprec); c.d2h(back, dq3)
    g5 = (3.0 * qk).reshape(4, 3, V, 2)[[2, 3, 0, 1]].reshape(12, V, 2)
    assert lu.rel_l2(back, g5) < 1e-7
    oprec = 4 if qprec == 8 else 8
    odt = np.float32 if qprec == 8 else np.float64
    dq4 = c.dev_malloc(qk.size * oprec)
    c.qkxtm_cast(dq4, oprec, dq, qprec)
    cast = np.empty(qk.shape, dtype=odt); c.d2h(cast, dq4)
    assert np.array_equal(cast, qk.astype(odt))
    # absorbVectorToDevice (lib/qudaQKXTM_Propagator.cpp:90-106)
    dprop = c.dev_malloc(144 * V * 2 * qprec)
    c.L.tmq_dev_memset(c.h, dprop, 0, 144 * V * 2 * qprec)
    c.qkxtm_absorb(dprop, dq, qprec, 2, 1)
    prop = np.empty((4, 4, 3, 3, V, 2), dtype=qdt); c.d2h(prop, dprop)
    assert np.array_equal(prop[:, 2, :, 1], qk.reshape(4, 3, V, 2))
    mask = np.ones((4, 4, 3, 3), bool); mask[:, 2, :, 1] = False
    assert not prop[mask].any()
    for p in (dq, dq3, dq4, dprop):
        c.dev_free(p)


def test_error_paths_fail_loudly(tmq):
    X = LATTICES[1]
    with pytest.raises(tmq.TmqError):
        tmq.Context((5, 4, 4, 4))                    # odd extent
    with pytest.raises(tmq.TmqError):
        tmq.Context(X, grid=(2, 1, 1, 1))            # x may not be partitioned
    c = tmq.Context(X)
    a, b = c.spinor(), c.spinor()
    with pytest.raises(tmq.TmqError):
        c.dslash(b, a, 0)                            # no gauge loaded
    with pytest.raises(tmq.TmqError):
        c.load_gauge(np.zeros((4, c.V, 3, 3, 2)), recon=8)
    c.load_gauge(lu.unit_gauge_qdp(X), t_boundary=1, recon=18)
    with pytest.raises(tmq.TmqError):
        c.matpc(b, a)                                # operator not set
    with pytest.raises(tmq.TmqError):
        c.dslash(a, a, 0)                            # aliasing
    c.set_op(KAPPA, MU, 0)
    f = c.spinor(8, tmq.FULL)
    with pytest.raises(tmq.TmqError):
        c.mdagm(f, f)                                # wrong subset
    s4 = c.spinor(4)
    with pytest.raises(tmq.TmqError):
        c.cg_mdagm(s4, s4)                           # solution must be fp64
    c.close()


# ---- 8-real gauge format (opt-in; csrc/tmq_site.cuh: reconstruct_from8, tools/recon8_study.py) ---------------------------------
def test_reconstruct_8_matches_oracle_and_the_12_real_path(tmq):
    """the trig-free 8-real link format: hop (both parities, daggers) and M^dag M against the oracle within the fp64 / fp32 bounds, the
    same CG iteration count as the CPU CG, the ghost-zone path (halo pack with U^dag from 8 reals), the clover build and the
    plaquette from the 8-real field; a unit field cannot be stored and is refused"""
    X = (4, 6, 4, 8)
    s = get_setup(tmq, X, 8); o = s.oracle(); c = s.ctx
    for prec in (8, 4):
        a, b = c.spinor(prec), c.spinor(prec)
        for out_parity in (0, 1):
            src = s.odd if out_parity == 0 else s.even
            a.set(src)
            for dagger in (0, 1):
                c.dslash(b, a, out_parity, dagger)
                assert lu.rel_l2(b.get(), o.dslash(s.gauge, src, out_parity, dagger)) < TOL[prec], (prec, out_parity, dagger)
        c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
        a.set(s.even)
        c.mdagm(b, a)
        assert lu.rel_l2(b.get(), o.mdagm(s.gauge, s.even, KAPPA, MU, 0)) < 2 * TOL[prec]
    assert abs(c.plaquette() - o.plaquette(s.gauge)) < 1e-13
    x, bb = c.spinor(8), c.spinor(8)
    bb.set(s.even)
    x_ref, it_ref, _, _ = o.cg_mdagm(s.gauge, s.even, KAPPA, MU, 0, tol=1e-10, maxiter=2000)
    info = c.cg_mdagm(x, bb, tol=1e-10, maxiter=2000)
    assert abs(info["iter"] - it_ref) <= 2 and info["true_res"] < 1.05e-10 and lu.rel_l2(x.get(), x_ref) < 1e-9
    info4 = c.cg_mdagm(x, bb, tol=1e-10, maxiter=2000, sloppy_prec=4)
    assert info4["true_res"] < 1.05e-10 and lu.rel_l2(x.get(), x_ref) < 1e-8
    # ghost-zone path and clover on the 8-real field
    c2 = tmq.Context(X)
    c2.force_partition((0, 0, 1, 1))
    c2.load_gauge(s.gauge, t_boundary=-1, recon=8)
    c2.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    a2, b2 = c2.spinor(8), c2.spinor(8)
    a2.set(s.even)
    c2.mdagm(b2, a2)
    assert lu.rel_l2(b2.get(), o.mdagm(s.gauge, s.even, KAPPA, MU, 0)) < 2e-13
    csw = 1.57551
    c2.clover_load(csw * KAPPA)
    o.set_clover(o.clover_compute(s.gauge, csw * KAPPA))
    c2.mdagm(b2, a2)
    want = o.mdagm(s.gauge, s.even, KAPPA, MU, 0)
    o.set_clover(None)
    assert lu.rel_l2(b2.get(), want) < 4e-13
    c2.close()
    # a unit field has U01 = U02 = 0 everywhere: refused, like any field the format cannot hold
    unit = np.zeros_like(s.gauge); unit[..., 0, 0, 0] = unit[..., 1, 1, 0] = unit[..., 2, 2, 0] = 1.0
    c3 = tmq.Context(X)
    with pytest.raises(tmq.TmqError):
        c3.load_gauge(unit, t_boundary=1, recon=8)
    c3.load_gauge(unit, t_boundary=1, recon=12)
    c3.close()
