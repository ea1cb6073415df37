"""GPU parity for BASELINE.json configs[1] and configs[2] (the other configs: 8^3x16 in test_gpu_parity.py / smoke(), 48^3x96 in
test_gpu_fullsize.py, 64^3x128 in bench.py's scale64 leg with its solution checksum):

  * 24^3 x 48, fp64 CG on M^dag M on a single B200 against the FULL CPU CG of the oracle: iteration count within +-2, true residual
    within 10 % of the CPU's and <= tol, solution relative L2 <= 10 tol (SURVEY.md 8d "parity gates");
  * 32^3 x 64, mixed precision (fp32 inner / fp64 outer) with the reference drivers' reliable_delta = 1e-4 (qkxtm/Calc_Loops.cpp:481):
    the same iteration count +-2 as the fp64 solve, fp64 true residual <= tol, the solution agrees with the fp64 one, and a single
    fp32 application of the hop / M^dag M agrees with the fp64 oracle to <= 1e-5 on that lattice.
  The 2-GPU T-sharded half of configs[2] is tests/sharded_parity.py (run under torchrun by tools/gpu_sharded.sh)."""
import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

KAPPA = 1.0 / (2.0 * (4.0 + 0.1))
MU = 0.1


def setup(X, recon=12):
    import tmq as T
    from oracle.oracle import Oracle
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    gauge = T.gen_gauge(X, seed=137, t_boundary=-1)
    Vh = int(np.prod(X)) // 2
    src = np.ascontiguousarray(T.gen_spinor(X, "z4", seed=100)[:Vh])
    c = T.Context(X)
    c.load_gauge(gauge, t_boundary=-1, recon=recon)
    c.set_op(KAPPA, MU, T.MATPC_EVEN_EVEN)
    return T, c, Oracle(X), gauge, src


def test_24x48_fp64_cg_against_the_full_cpu_cg():
    X = (24, 24, 24, 48)
    tol = 1e-9
    T, c, o, gauge, src = setup(X)
    x_ref, it_ref, tr_ref, hist_ref = o.cg_mdagm(gauge, src, KAPPA, MU, 0, tol=tol, maxiter=5000)
    b, x = c.spinor(), c.spinor()
    b.set(src)
    info = c.cg_mdagm(x, b, tol=tol, maxiter=5000)
    assert abs(info["iter"] - it_ref) <= 2, (info["iter"], it_ref)
    assert info["true_res"] <= 1.05 * tol
    assert abs(info["true_res"] - tr_ref) <= 0.1 * tr_ref, (info["true_res"], tr_ref)
    assert lu.rel_l2(x.get(), x_ref) <= 10 * tol
    hist = c.cg_history(it_ref + 1)
    n = min(info["iter"], it_ref)
    assert np.allclose(hist[: n + 1], hist_ref[: n + 1], rtol=1e-7, atol=0)
    # single applications on this lattice, fp64: <= 1e-13
    a, y = c.spinor(), c.spinor()
    a.set(src)
    c.dslash(y, a, 1, 0)
    assert lu.rel_l2(y.get(), o.dslash(gauge, src, 1, 0)) < 1e-13
    c.mdagm(y, a)
    assert lu.rel_l2(y.get(), o.mdagm(gauge, src, KAPPA, MU, 0)) < 2e-13
    c.close()


def test_32x64_mixed_precision_cg_at_the_drivers_reliable_delta():
    X = (32, 32, 32, 64)
    tol = 1e-9
    T, c, o, gauge, src = setup(X)
    b, x, xm = c.spinor(), c.spinor(), c.spinor()
    b.set(src)
    i64 = c.cg_mdagm(x, b, tol=tol, maxiter=5000)
    im = c.cg_mdagm(xm, b, tol=tol, maxiter=5000, reliable_delta=1e-4, sloppy_prec=4)      # qkxtm/Calc_Loops.cpp:481
    assert abs(im["iter"] - i64["iter"]) <= 2, (im["iter"], i64["iter"])
    assert im["true_res"] <= 1.05 * tol and i64["true_res"] <= 1.05 * tol
    assert lu.rel_l2(xm.get(), x.get()) <= 10 * tol
    # the true residual, recomputed here through the fp64 operator
    r = c.spinor()
    c.mdagm(r, xm)
    c.axpy(-1.0, b, r)
    assert np.sqrt(c.norm2(r) / c.norm2(b)) <= 1.05 * tol
    # the CPU CG takes the same number of iterations (full CPU solve: about 0.3 s per iteration on 16 cores)
    _, it_ref, tr_ref, _ = o.cg_mdagm(gauge, src, KAPPA, MU, 0, tol=tol, maxiter=5000)
    assert abs(i64["iter"] - it_ref) <= 2 and abs(im["iter"] - it_ref) <= 2
    assert abs(i64["true_res"] - tr_ref) <= 0.1 * tr_ref
    # fp32 single applications against the fp64 oracle: <= 1e-5
    a4, y4 = c.spinor(4), c.spinor(4)
    a4.set(src)
    c.dslash(y4, a4, 1, 0)
    assert lu.rel_l2(y4.get(), o.dslash(gauge, src, 1, 0)) < 1e-5
    c.dslash(y4, a4, 1, 1)
    assert lu.rel_l2(y4.get(), o.dslash(gauge, src, 1, 1)) < 1e-5
    c.mdagm(y4, a4)
    assert lu.rel_l2(y4.get(), o.mdagm(gauge, src, KAPPA, MU, 0)) < 1e-5
    c.close()
