"""GPU tests of the drop-in claim itself: the REFERENCE'S OWN drivers (qkxtm/Calc_Loops.cpp, qkxtm/MG_Bench.cpp with qkxtm/QKXTM_util.cpp and
qkxtm/misc.cpp, compiled unmodified from /root/reference by oracle/Makefile into oracle/_ref/dropin/ and linked against libqkxtm_tmq.so +
libtmq.so) run on the GPU through their own command lines, read an ILDG / LIME configuration with the reference's own reader, and what
they compute is checked against the CPU oracle.  Also calc_loops (include/qudaQKXTM.h:501-507) through this repository's own test driver
with spin-colour dilution and hierarchical probing."""
import os
import re
import subprocess

import numpy as np
import pytest

import lattice_util as lu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
DRV = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "qkxtm_invert_test")
X = (4, 6, 4, 8)
V = int(np.prod(X))
KAPPA = 0.1219512195
MU = 0.1


def _cplx(a):
    return np.ascontiguousarray(a[..., 0] + 1j * a[..., 1]).ravel()


def _real(v, n):
    return np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(n, 4, 3, 2))


def parse_dump(path, nrec_max=10 ** 6):
    """records of QKXTM_LOOP_DUMP / the test driver's hook: 8 doubles, [source], vector (plug-in AoS order, x lexicographic)"""
    raw = np.fromfile(path, dtype=np.float64)
    recs, p = [], 0
    while p < raw.size and len(recs) < nrec_max:
        h = raw[p:p + 8]; p += 8
        r = dict(kind=int(h[0]), isrc=int(h[1]), ih=int(h[2]), sc=int(h[3]), dstep=int(h[4]), nev_defl=int(h[5]), val=float(h[6]), true_res=float(h[7]))
        if r["kind"] == 1:
            r["src"] = raw[p:p + V * 24].reshape(V, 4, 3, 2); p += V * 24
        r["x"] = raw[p:p + V * 24].reshape(V, 4, 3, 2); p += V * 24
        recs.append(r)
    assert p == raw.size
    return recs


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    import tmq
    from oracle.oracle import Oracle
    d = tmp_path_factory.mktemp("dropin")
    raw = tmq.gen_gauge(X, seed=137, t_boundary=+1)                   # what is on disk: no boundary condition applied
    conf = str(d / "conf.lime")
    tmq.lime_write_gauge(conf, raw, X, kappa=KAPPA, mu=MU)
    gauge = raw.copy()
    tmq.apply_t_boundary(gauge, X, t_boundary=-1)                     # what the drivers hand to loadGaugeQuda (applyBoundaryCondition)
    assert np.array_equal(gauge, tmq.gen_gauge(X, seed=137, t_boundary=-1))
    return dict(tmq=tmq, o=Oracle(X), conf=conf, gauge=gauge, dir=d)


def run_ref(exe, args, env_extra=None, timeout=600):
    path = os.path.join(DROPIN, exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/dropin/%s not built (needs /root/reference at build time)" % exe)
    e = dict(os.environ)
    e.update(env_extra or {})
    p = subprocess.run([path] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout, env=e)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout, p.stderr


COMMON = ["--dim", *X, "--prec", "double", "--prec-sloppy", "double", "--recon", "12", "--recon-sloppy", "12", "--dslash-type", "twisted-mass",
          "--kappa", KAPPA, "--mu", MU, "--inv-type", "cg", "--solve-type", "normop-pc", "--tol", "1e-9", "--niter", "4000",
          "--mass-normalization", "kappa", "--verbosity-level", "summarize"]


def test_reference_Calc_Loops_driver_on_libtmq(env):
    """qkxtm/Calc_Loops.cpp main() -> readLimeGauge (the reference's reader) -> applyBoundaryCondition -> initQuda -> init_qudaQKXTM ->
    loadGaugeQuda -> calc_loops.  Checked: the noise vectors are the GSL-ranlux Z4 stream of the given seed; every solve takes the CPU
    CG's iteration count (+-2) and solves M x = source; the low modes are ARPACK's on the oracle operator; the projected solutions
    are orthogonal to them."""
    tmq, o, gauge = env["tmq"], env["o"], env["gauge"]
    from oracle.oracle import eigs_reference
    seed, nstoch, nev = 4711, 2, 4
    n = 24 * o.Vh

    def A(v):
        y = o.mat(gauge, _real(v, 2 * o.Vh), KAPPA, -MU, 0)
        return _cplx(o.mat(gauge, y, KAPPA, -MU, 1))
    lam_ref, U_ref = eigs_reference(A, n, nev, 30, "SR", tol=1e-12)
    dump = str(env["dir"] / "loops.bin")
    out, err = run_ref("Calc_Loops", COMMON + ["--load-gauge", env["conf"], "--Nstoch", nstoch, "--NdumpStep", 1, "--seed", seed, "--useEven", "true",
                                               "--nEv", nev, "--nKv", 30, "--PolyDeg", 24, "--isACC", "true", "--aminARPACK", 1.6 * lam_ref[-1],
                                               "--amaxARPACK", 3.5, "--tolARPACK", "1e-12", "--maxIterARPACK", 400, "--useFullOp", "true",
                                               "--spectrumPart", "SR", "--defl-steps", 2, "--defl-step-NeV", 0, 0, "--defl-step-NeV", 1, nev,
                                               "--loop-file-format", "ASCII", "--Q-sqMax", 0],
                       env_extra={"QKXTM_LOOP_DUMP": dump})
    assert "### calc_loops: Loop calculation begins now" in out and "Stochastic part calculation Done" in out
    plaq = float(re.search(r"Calculated plaquette in double precision is (\S+)", out).group(1))
    assert abs(plaq - o.plaquette(gauge)) < 1e-6
    iters = [int(m) for m in re.findall(r"CG: Convergence at (\d+) iterations", out)]
    assert len(iters) == nstoch
    recs = parse_dump(dump)
    evs = [r for r in recs if r["kind"] == 0]
    sol = [r for r in recs if r["kind"] == 1]
    assert len(evs) == nev and len(sol) == nstoch * 2
    # exact part: eigenvalues and eigenvectors of M_full^dag M_full (mu < 0: "for the loops we invert the negative mu")
    assert np.allclose([r["val"] for r in evs], lam_ref, rtol=1e-9)
    U = np.stack([_cplx(lu.spinor_eo_from_lex(r["x"], X)) for r in evs], axis=1)
    assert np.abs(U.conj().T @ U - np.eye(nev)).max() < 1e-10
    assert np.linalg.svd(U_ref.conj().T @ U, compute_uv=False).min() > 1 - 1e-8
    # stochastic part
    rng = tmq.Ranlux(seed)                                            # rank 0: seed + 0 * seed
    for isrc in range(nstoch):
        noise = rng.z4(V * 12).reshape(V, 4, 3, 2)
        r0, r1 = sol[2 * isrc], sol[2 * isrc + 1]
        assert (r0["isrc"], r0["dstep"], r0["nev_defl"], r1["dstep"], r1["nev_defl"]) == (isrc, 0, 0, 1, nev)
        assert np.array_equal(r0["src"], noise)                       # bit-identical to the reference's gsl_rng_ranlux stream
        b_eo = lu.spinor_eo_from_lex(noise, X)
        x_eo = lu.spinor_eo_from_lex(r0["x"], X)
        assert lu.rel_l2(o.mat(gauge, x_eo, KAPPA, -MU, 0), b_eo) < 1e-7
        # iteration count of the CPU CG on the same system (even-even, prepared source, M^dag applied first)
        src_pc = o.prepare(gauge, b_eo, KAPPA, -MU, 0)
        rhs = o.matpc(gauge, src_pc, KAPPA, -MU, 0, 1)
        _, it_ref, tr_ref, _ = o.cg_mdagm(gauge, rhs, KAPPA, -MU, 0, tol=1e-9, maxiter=4000)
        assert abs(iters[isrc] - it_ref) <= 2 and int(r0["val"]) == iters[isrc]
        assert r0["true_res"] <= 1.05e-9
        # projection: x1 = (1 - U U^dag) x0
        xp = _cplx(lu.spinor_eo_from_lex(r1["x"], X))
        x0 = _cplx(x_eo)
        assert lu.rel_l2(xp, x0 - U @ (U.conj().T @ x0)) < 1e-10
        assert np.abs(U.conj().T @ xp).max() < 1e-9 * np.linalg.norm(x0)


def test_reference_MG_Bench_driver_on_libtmq(env):
    """qkxtm/MG_Bench.cpp main(): reads the configuration twice (links + "smeared" links for the plaquette print), newMultigridQuda (inert
    here), then MG_bench with the reference's exact signature: 12 point-source solves.  The iteration counts are the CPU CG's."""
    o, gauge = env["o"], env["gauge"]
    out, err = run_ref("MG_Bench", COMMON + ["--load-gauge", env["conf"], "--load-gauge-smeared", env["conf"], "--nsmearGauss", 0, "--useEven", "true"])
    assert "Begin MG bench routine" in out
    assert "multigrid preconditioner is not provided" in err and "GCR (+ multigrid) was requested" in err      # the driver hard-wires GCR + MG
    assert out.count("Inversion up =") == 12
    iters = [int(m) for m in re.findall(r"CG: Convergence at (\d+) iterations", out)]
    assert len(iters) == 12
    for isc in (0, 7):
        b = np.zeros((V, 4, 3, 2)); b.reshape(V, 12, 2)[0, isc, 0] = 1.0
        b_eo = lu.spinor_eo_from_lex(b, X)
        rhs = o.matpc(gauge, o.prepare(gauge, b_eo, KAPPA, MU, 0), KAPPA, MU, 0, 1)
        _, it_ref, _, _ = o.cg_mdagm(gauge, rhs, KAPPA, MU, 0, tol=1e-9, maxiter=4000)
        assert abs(iters[isc] - it_ref) <= 2, (isc, iters[isc], it_ref)


def test_reference_driver_refuses_what_is_not_built(env):
    """--prec single asks for an fp32 solution field, which the QKXTM containers cannot take (the upload kernel writes double2,
    lib/qudaQKXTM_kernels.cu:1031): refused with an errorQuda abort, not silently replaced"""
    path = os.path.join(DROPIN, "MG_Bench")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/dropin not built")
    args = [str(a) for a in COMMON]
    args[args.index("--prec") + 1] = "single"
    p = subprocess.run([path] + args + ["--load-gauge", env["conf"], "--load-gauge-smeared", env["conf"], "--nsmearGauss", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "ERROR" in p.stderr


@pytest.mark.parametrize("mode", ["dilution", "probing", "unity"])
def test_calc_loops_dilution_and_probing(env, tmp_path, mode):
    """calc_loops through this repository's driver (hook installed): spin-colour dilution (12 solves per noise vector), hierarchical
    probing (k = 1: 2 Hadamard vectors) and unity sources; every diluted source is what the reference's routines give (pinned on the CPU
    by tests/test_calc_loops_noise.py) and every solution solves M x = diluted source."""
    tmq, o, gauge = env["tmq"], env["o"], env["gauge"]
    outf = str(tmp_path / "out.bin")
    flags = {"dilution": ["--spinColorDil", "yes"], "probing": ["--k-probing", "1"], "unity": ["--source-type", "unity"]}[mode]
    cmd = [DRV, "--dim", *[str(x) for x in X], "--test", "calcloops", "--tol", "1e-9", "--recon", "12", "--kappa", str(KAPPA), "--mu", str(MU),
           "--Nstoch", "1", "--seed", "99", "--nEv", "0", "--defl-steps", "1", "0", "--out", outf] + flags
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    recs = parse_dump(outf)
    noise = tmq.Ranlux(99).z4(V * 12, unity=(mode == "unity")).reshape(V, 12, 2)
    if mode == "dilution":
        want = []
        for sc in range(12):
            w = np.zeros_like(noise); w[:, sc] = noise[:, sc]; want.append(w)
    elif mode == "probing":
        Vc = tmq.hch_coloring(X, 1, 4)
        want = [np.array([tmq.hadamard_element(int(c), ih) for c in Vc], dtype=np.float64)[:, None, None] * noise for ih in range(2)]
    else:
        want = [noise]
        assert np.all(noise[..., 0] == 1) and np.all(noise[..., 1] == 0)
    assert len(recs) == len(want)
    for r, w in zip(recs, want):
        assert np.array_equal(r["src"].reshape(V, 12, 2), w)
        b_eo = lu.spinor_eo_from_lex(np.ascontiguousarray(w.reshape(V, 4, 3, 2)), X)
        assert lu.rel_l2(o.mat(gauge, lu.spinor_eo_from_lex(r["x"], X), KAPPA, -MU, 0), b_eo) < 1e-7
