"""GPU tests of the C++ QKXTM host layer (include/qudaQKXTM_tmq.h, host/qudaQKXTM_tmq.cpp) through the
qkxtm_invert_test driver: the reference-shaped call sequences (initQuda -> init_qudaQKXTM -> loadGaugeQuda ->
invertQuda / MG_bench / calc_loops solve / ApplyMdagM) checked against the CPU oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "qkxtm_invert_test")
X = (4, 6, 4, 8)
KAPPA = 1.0 / (2.0 * 4.1)
MU = 0.1


def run(tmp_path, *flags):
    out = str(tmp_path / "out.bin")
    cmd = [DRV, "--dim"] + [str(x) for x in X] + list(flags) + ["--out", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    m = re.search(r"RESULT .*iter=(\d+) true_res=(\S+)", p.stdout)
    return np.fromfile(out, dtype=np.float64), int(m.group(1)), float(m.group(2)), p.stdout


@pytest.fixture(scope="module")
def env():
    import tmq
    from oracle.oracle import Oracle
    o = Oracle(X)
    return o, tmq.gen_gauge(X, seed=137, t_boundary=-1), tmq


@pytest.mark.parametrize("flags", [[], ["--recon", "12"], ["--prec-sloppy", "single", "--recon", "12"], ["--matpc", "odd-odd"]])
def test_invertQuda_solves_full_system(tmp_path, env, flags):
    o, gauge, tmq = env
    x, it, tr, _ = run(tmp_path, "--test", "invert", "--tol", "1e-10", *flags)
    b = tmq.gen_spinor(X, "z4", seed=100)
    assert tr <= 1.05e-10 and it > 5
    assert lu.rel_l2(o.mat(gauge, x.reshape(b.shape), KAPPA, MU, 0), b) < 1e-8


def test_mass_normalization_rescales_by_2kappa(tmp_path, env):
    o, gauge, tmq = env
    xk, _, _, _ = run(tmp_path, "--test", "invert", "--tol", "1e-10")
    xm, _, _, _ = run(tmp_path, "--test", "invert", "--tol", "1e-10", "--mass-normalization", "mass")
    assert lu.rel_l2(xm, 2 * KAPPA * xk) < 1e-12          # lib/qudaQKXTM_interface.cpp:200-203


def test_calc_loops_solve_in_plugin_host_order(tmp_path, env):
    o, gauge, tmq = env
    x, it, tr, _ = run(tmp_path, "--test", "loops", "--tol", "1e-9", "--mu", "-0.1", "--recon", "12")   # loops use mu < 0 (Calc_Loops.cpp:424)
    b_lex = tmq.gen_spinor(X, "z4", seed=100, eo_order=False)
    x_eo = lu.spinor_eo_from_lex(x.reshape(b_lex.shape), X)
    assert tr <= 1.05e-9
    assert lu.rel_l2(o.mat(gauge, x_eo, KAPPA, -MU, 0), lu.spinor_eo_from_lex(b_lex, X)) < 1e-7


def test_ApplyMdagM_and_MatQuda(tmp_path, env):
    o, gauge, tmq = env
    y, _, _, _ = run(tmp_path, "--test", "mdagm", "--source", "gaussian", "--seed", "101")
    src = lu.spinor_eo_from_lex(tmq.gen_spinor(X, "gaussian", seed=101, eo_order=False), X)
    y_eo = lu.spinor_eo_from_lex(y.reshape(src.shape), X)
    assert lu.rel_l2(y_eo[: o.Vh], o.mdagm(gauge, np.ascontiguousarray(src[: o.Vh]), KAPPA, MU, 0)) < 2e-13
    assert not y_eo[o.Vh:].any()                            # absent parity zero-filled (downloadFromCuda_core.h)
    m, _, _, _ = run(tmp_path, "--test", "mat", "--source", "gaussian", "--seed", "101")
    src_eo = tmq.gen_spinor(X, "gaussian", seed=101)
    assert lu.rel_l2(m.reshape(src_eo.shape), o.mat(gauge, src_eo, KAPPA, MU, 0)) < 1e-13


def test_MG_bench_twelve_columns_and_plaquette(tmp_path, env):
    o, gauge, tmq = env
    prop, it, tr, log = run(tmp_path, "--test", "mgbench", "--tol", "1e-9", "--recon", "12")
    V = o.V
    prop = prop.reshape(12, V, 4, 3, 2)
    plaq = float(re.search(r"Calculated plaquette in double precision is (\S+)", log).group(1))
    assert abs(plaq - o.plaquette(gauge)) < 1e-6            # printed with %lf
    assert log.count("Inversion up =") == 12
    for isc in (0, 5, 11):
        b = np.zeros((V, 4, 3, 2)); b.reshape(V, 12, 2)[0, isc, 0] = 1.0
        x_eo = lu.spinor_eo_from_lex(prop[isc], X)
        assert lu.rel_l2(o.mat(gauge, x_eo, KAPPA, MU, 0), lu.spinor_eo_from_lex(b, X)) < 1e-7


def test_driver_rejects_unsupported_parameters(tmp_path):
    p = subprocess.run([DRV, "--dim", "5", "4", "4", "4"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "ERROR" in p.stderr      # errorQuda aborts


def test_deflation_class_eigensolver_and_deflate(tmp_path, env):
    """QKXTM_Deflation (include/qudaQKXTM.h:391-475): eigenSolver on the even-even asymmetric M^dag M (the operator the
    reference's eigensolver uses, qkxtm/Calc_Loops.cpp:712-713), copyEigenVectorToQKXTM_Vector and deflateVector"""
    o, gauge, tmq = env
    nev = 4
    r, _, _, out = run(tmp_path, "--test", "eig", "--matpc", "even-even-asym", "--nEv", str(nev), "--nKv", "24", "--PolyDeg", "20",
                       "--amin", "0.34", "--amax", "2.0", "--tolArpack", "1e-11", "--source", "gaussian")
    V = int(np.prod(X))
    ev, rs = r[:nev], r[nev:2 * nev]
    v0 = lu.spinor_eo_from_lex(r[2 * nev: 2 * nev + V * 24].reshape(V, 4, 3, 2), X)
    xd = lu.spinor_eo_from_lex(r[2 * nev + V * 24:].reshape(V, 4, 3, 2), X)
    Vh = V // 2
    assert np.all(np.diff(ev) >= 0) and np.all(rs < 1e-8), (ev, rs)
    # eigenvector 0: even parity normalised, odd parity zero, eigen-equation holds with the oracle operator
    e0 = np.ascontiguousarray(v0[:Vh])
    assert np.all(v0[Vh:] == 0) and abs(np.sum(e0 * e0) - 1) < 1e-10
    Ae = o.mdagm(gauge, e0, KAPPA, MU, 2)
    assert lu.rel_l2(Ae, ev[0] * e0) < 1e-8
    # smallest eigenvalue agrees with ARPACK on the oracle operator
    from oracle.oracle import eigs_reference
    cplx = lambda a: np.ascontiguousarray(a[..., 0] + 1j * a[..., 1]).ravel()
    real = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(Vh, 4, 3, 2))
    lam, U = eigs_reference(lambda v: cplx(o.mdagm(gauge, real(v), KAPPA, MU, 2)), 12 * Vh, nev, 24, "SR", tol=1e-12)
    assert np.allclose(ev, lam, rtol=1e-9), (ev, lam)
    # deflated source = U Lambda^-1 U^dag b_even, odd parity zero
    b = lu.spinor_eo_from_lex(tmq.gen_spinor(X, "gaussian", seed=100, eo_order=False), X)
    ref = U @ ((U.conj().T @ cplx(np.ascontiguousarray(b[:Vh]))) / lam)
    assert np.all(xd[Vh:] == 0) and lu.rel_l2(cplx(np.ascontiguousarray(xd[:Vh])), ref) < 1e-7


def test_twisted_clover_invert_through_the_shim(tmp_path, env):
    """--dslash-type twisted-clover: loadCloverQuda(NULL, NULL, &inv_param) after loadGaugeQuda, then invertQuda
    (qkxtm/MG_Bench.cpp:598-608); the solution satisfies the oracle's full twisted-clover operator"""
    o, gauge, tmq = env
    csw = 1.57551
    x, it, tr, _ = run(tmp_path, "--test", "invert", "--tol", "1e-10", "--dslash-type", "twisted-clover", "--csw", str(csw), "--recon", "12")
    b = tmq.gen_spinor(X, "z4", seed=100)
    from oracle.oracle import Oracle
    oc = Oracle(X)
    oc.set_clover(oc.clover_compute(gauge, csw * KAPPA))
    r = oc.mat(gauge, x.reshape(b.shape), KAPPA, MU, 0)
    oc.set_clover(None)
    assert tr <= 1.05e-10 and it > 5
    assert lu.rel_l2(r, b) < 1e-8


def test_MG_bench_with_gaussian_smeared_sources(tmp_path, env):
    """MG_bench with nsmearGauss > 0 (lib/qudaQKXTM_interface.cpp:181-185): every point source is Gaussian-smeared with the
    links passed as gaugeSmeared before the solve, so M x = S delta; S from the oracle's restatement of Gauss_core.h (pinned
    to the reference's kernel body by tests/test_ref_kernels.py).  The driver also runs testGaussSmearing, whose 12 printed
    numbers are S delta at the origin."""
    o, gauge, tmq = env
    from oracle.oracle import gauss_smear
    nsm, alpha = 3, 4.0
    r, _, _, out = run(tmp_path, "--test", "mgbench", "--tol", "1e-10", "--nsmearGauss", str(nsm), "--alphaGauss", str(alpha), "--recon", "12")
    V = int(np.prod(X))
    cols = r.reshape(12, V, 4, 3, 2)
    # the links of the synthetic configuration in the QKXTM layout [dir][c1][c2][x_lex] (anti-periodic sign included, as the
    # driver passes the same field for both arguments)
    U = lu.r2c(np.stack([lu.spinor_lex_from_eo(gauge[mu], X) for mu in range(4)]))      # [4][x_lex][3][3]
    Uq = np.transpose(U, (0, 2, 3, 1))
    for isc in (0, 7):
        delta = np.zeros((12, V), dtype=np.complex128); delta[isc, 0] = 1.0
        Sd = gauss_smear(delta, Uq, X, alpha, nsm)                                  # [12][V]
        b_lex = lu.c2r(np.transpose(Sd.reshape(4, 3, V), (2, 0, 1)))               # [x_lex][s][c][ri]
        b_eo = lu.spinor_eo_from_lex(b_lex, X)
        x_eo = lu.spinor_eo_from_lex(cols[isc], X)
        assert lu.rel_l2(o.mat(gauge, x_eo, KAPPA, MU, 0), b_eo) < 1e-8, isc
        if isc == 0:
            printed = np.array([[float(v) for v in ln.split()] for ln in out.splitlines()
                                if len(ln.split()) == 2 and ln.lstrip()[0] in "+-" and "e" in ln][:12])
            assert printed.shape == (12, 2) and np.allclose(printed, b_lex[0].reshape(12, 2), rtol=1e-6, atol=1e-12)


def test_calcLowModeProjection_entry_point(tmp_path, env):
    """calcLowModeProjection (lib/qudaQKXTM_interface.cpp:1342-1401): parameter checks + QKXTM_Deflation::eigenSolver; the eigenvalues
    agree with ARPACK on the oracle operator, and a symmetric operator type is refused like in the reference"""
    o, gauge, tmq = env
    from oracle.oracle import eigs_reference
    nev = 4
    r, nconv, _, _ = run(tmp_path, "--test", "lowmodes", "--matpc", "even-even-asym", "--nEv", str(nev), "--nKv", "24", "--PolyDeg", "20",
                         "--amin", "0.34", "--amax", "2.0", "--tolArpack", "1e-11")
    assert nconv == nev
    cplx = lambda v: np.ascontiguousarray(v[..., 0] + 1j * v[..., 1]).ravel()
    real = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(o.Vh, 4, 3, 2))
    A = lambda v: cplx(o.mdagm(gauge, real(v), KAPPA, MU, 2))
    lam, _ = eigs_reference(A, 12 * o.Vh, nev, 24, "SR", poly=(20, 0.34, 2.0), tol=1e-12)
    assert np.allclose(r[:nev], lam, rtol=1e-9)
    p = subprocess.run([DRV, "--dim"] + [str(x) for x in X] + ["--test", "lowmodes", "--matpc", "even-even"], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "Only asymmetric operators are supported in deflation" in p.stderr


def test_invertMultiSrcQuda_pipeline_and_e2e_driver(tmp_path, env):
    """--test e2e: invertQuda (fp64, then fp32-sloppy) and invertMultiSrcQuda with page-locked host fields; the uploads / downloads of
    neighbouring columns run behind each solve (tmq_host_prefetch / tmq_spinor_to_host_async).  The last column's solution must solve
    M x = b for ITS source, i.e. no column was mixed up by the slot reuse, and the mass normalisation is applied on the device."""
    import json
    o, gauge, tmq = env
    nsrc = 5
    for norm, scale in (("kappa", 1.0), ("mass", 2 * KAPPA)):
        x, _, _, out = run(tmp_path, "--test", "e2e", "--tol", "1e-10", "--recon", "12", "--nsrc", str(nsrc), "--seed", "100", "--mass-normalization", norm)
        r = json.loads(re.search(r"RESULT_E2E (\{.*\})", out).group(1))
        b = tmq.gen_spinor(X, "z4", seed=100 + ((nsrc - 1) & 1))          # the driver alternates two sources (seed, seed + 1)
        assert lu.rel_l2(o.mat(gauge, x.reshape(b.shape) / scale, KAPPA, MU, 0), b) < 1e-8
        assert r["single_true_res"] <= 1.05e-10 and r["mixed_true_res"] <= 1.05e-10 and r["multi_true_res"] <= 1.05e-10
        assert abs(r["single_iter"] - r["mixed_iter"]) <= 2
        assert r["nsrc"] == nsrc and abs(r["multi_iter"] - nsrc * r["single_iter"]) <= 2 * nsrc
