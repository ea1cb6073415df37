"""GPU parity tests of the step after the solves: propagator container kernels and the meson two-point contraction
(csrc/tmq_contract.cu) against the REFERENCE'S OWN kernel bodies -- the golden fixture made from contractMesons_core.h (with the
reference's channel tables), rotateToPhysicalBase_core.h, apply_gamma5_propagator_core.h, conjugate_*_core.h compiled for the
CPU (tests/golden/make_golden_contract.py) -- and, on other shapes, against the oracle's numpy restatement (pinned to the same
kernels by tests/test_ref_contract.py).  Tolerances: fp64 propagators 1e-12 of the largest entry (the sums run in a different
order), fp32 propagators 1e-5 (the reference accumulates in float, the device kernel multiplies in float and sums in double)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import lattice_util as lu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_contract as G  # noqa: E402
from test_gpu_smear import Dev, _c, tmq  # noqa: E402,F401

from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "qkxtm_invert_test")


def _relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_meson_contraction_matches_reference_fixture(tmq):
    gold = np.load(G.FIXTURE)
    p1, p2 = G.contract_inputs()
    d = Dev(tmq, G.X)
    moms = G.momenta()
    # fp64 propagators: the double instantiation of the reference's kernel body
    d1, d2 = d.put(p1), d.put(p2)
    mom, pos = d.c.qkxtm_contract_mesons(d1, d2, 8, moms, G.SRC, global_T=G.X[3], pos=True)
    assert _relmax(mom, _c(gold["mom_double"])) < 1e-13
    # fp32 propagators: what the reference launches
    f1, f2 = d.put(p1.astype(np.float32)), d.put(p2.astype(np.float32))
    momf, posf = d.c.qkxtm_contract_mesons(f1, f2, 4, moms, G.SRC, global_T=G.X[3], pos=True)
    assert _relmax(momf, _c(gold["mom_float"]).astype(np.complex128)) < 1e-5
    want_pos = _c(gold["pos_float"]).astype(np.complex128).reshape(-1, 2, 10)      # [T][V3] = x_lex
    assert _relmax(posf, want_pos) < 1e-5
    assert _relmax(pos, want_pos) < 1e-5
    # repeatable to the bit (fixed summation order)
    mom2, _ = d.c.qkxtm_contract_mesons(d1, d2, 8, moms, G.SRC, global_T=G.X[3])
    assert np.array_equal(mom, mom2)
    d.close()


@pytest.mark.parametrize("X,q_sq,src", [((6, 4, 10, 4), 4, (5, 0, 7)), ((16, 12, 8, 6), 2, (3, 11, 2)), ((36, 4, 34, 2), 1, (0, 1, 33))])
def test_meson_contraction_other_shapes_vs_restatement(tmq, X, q_sq, src):
    """extents that are not multiples of the warp size, rows longer than a warp, more momenta"""
    rng = np.random.default_rng(17)
    V = int(np.prod(X))
    p1 = rng.standard_normal((4, 4, 3, 3, V, 2)); p2 = rng.standard_normal((4, 4, 3, 3, V, 2))
    moms = O.create_momenta(q_sq)
    want = O.contract_mesons_mom(_c(p1), _c(p2), X, moms, src)
    d = Dev(tmq, X)
    mom, pos = d.c.qkxtm_contract_mesons(d.put(p1), d.put(p2), 8, moms, src, global_T=X[3], pos=True)
    assert _relmax(mom, want) < 1e-12
    site = np.stack([O.contract_mesons_site(_c(p1)), O.contract_mesons_site(_c(p2))])       # [2][10][V]
    assert _relmax(pos, np.transpose(site, (2, 0, 1))) < 1e-13
    # a momentum list in another order / with repeats maps to the same numbers
    perm = [moms[i] for i in (3, 0, 3, len(moms) - 1)]
    mom_p, _ = d.c.qkxtm_contract_mesons(d.bufs[0], d.bufs[1], 8, perm, src, global_T=X[3])
    assert np.array_equal(mom_p[:, 0], mom[:, 3]) and np.array_equal(mom_p[:, 1], mom[:, 0]) and np.array_equal(mom_p[:, 3], mom[:, -1])
    d.close()


def test_baryon_contraction_matches_reference_fixture(tmq):
    """ten baryon channels x 4x4 spin x two propagator assignments against the reference's kernel body (contractBaryons_core.h with
    the reference's tables) compiled for the CPU: its double instantiation to 1e-12, the float one it launches to 2e-5"""
    gold = np.load(G.FIXTURE)
    p1, p2 = G.contract_inputs()
    d = Dev(tmq, G.X)
    moms = G.baryon_momenta()
    got = d.c.qkxtm_contract_baryons(d.put(p1), d.put(p2), 8, moms, G.SRC, G.X[3])
    want = _c(gold["baryon_mom_double"])
    assert got.shape == want.shape
    for ip in range(10):
        for iu in range(2):
            assert _relmax(got[:, :, iu, ip], want[:, :, iu, ip]) < 1e-12, (ip, iu)
    gotf = d.c.qkxtm_contract_baryons(d.put(p1.astype(np.float32)), d.put(p2.astype(np.float32)), 4, moms, G.SRC, G.X[3])
    wantf = _c(gold["baryon_mom_float"]).astype(np.complex128)
    for ip in range(10):
        assert _relmax(gotf[:, :, :, ip], wantf[:, :, :, ip]) < 2e-5, ip
        assert _relmax(gotf[:, :, :, ip], want[:, :, :, ip]) < 2e-5, ip
    got2 = d.c.qkxtm_contract_baryons(d.bufs[0], d.bufs[1], 8, moms, G.SRC, G.X[3])
    assert np.array_equal(got, got2)                                  # fixed summation order
    # large lattices are contracted in passes of a few time slices (2 GiB of site values per pass): force 1 and 4 slices per pass
    for nslices in (1, 4):
        d.c.set_option(tmq.OPT_CONTRACT_SLICES, nslices)
        assert np.array_equal(d.c.qkxtm_contract_baryons(d.bufs[0], d.bufs[1], 8, moms, G.SRC, G.X[3]), got), nslices
    d.c.set_option(tmq.OPT_CONTRACT_SLICES, 0)
    d.close()


@pytest.mark.parametrize("X,src", [((6, 4, 2, 4), (5, 0, 1)), ((10, 2, 6, 2), (3, 1, 2))])
def test_baryon_contraction_other_shapes_vs_live_reference_kernel(tmq, X, src):
    """sites per time slice that are not a multiple of the 32-site CTA tile, against the reference kernel body run live (the prebuilt
    oracle/_ref library travels to the GPU box)"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(29)
    V = int(np.prod(X))
    p1 = rng.standard_normal((4, 4, 3, 3, V, 2)); p2 = rng.standard_normal((4, 4, 3, 3, V, 2))
    moms = [(0, 0, 0), (1, -1, 0), (0, 0, 2)]
    want = _c(ref.Ref(X).contract_baryons_mom(p1, p2, moms, src))
    d = Dev(tmq, X)
    got = d.c.qkxtm_contract_baryons(d.put(p1), d.put(p2), 8, moms, src, X[3])
    for ip in range(10):
        assert _relmax(got[:, :, :, ip], want[:, :, :, ip]) < 1e-12, ip
    d.close()


def test_sequential_sources_and_local_insertion_match_reference_fixture(tmq):
    """seqSourceFixSinkPart1 / Part2 and the ultra-local fixed-sink contraction against the reference's kernel bodies (fixture), plus
    every projector / particle combination against the restatement (pinned to the same kernels by tests/test_ref_contract.py)"""
    gold = np.load(G.FIXTURE)
    p1, p2 = G.contract_inputs()
    X = G.X
    V = int(np.prod(X)); V3 = V // X[3]
    d = Dev(tmq, X)
    t1, t2 = np.ascontiguousarray(p1[..., 2 * V3:3 * V3, :]), np.ascontiguousarray(p2[..., 2 * V3:3 * V3, :])
    d1, d2 = d.put(t1), d.put(t2)
    ts = 4
    for key, (part, pid, particle, nu, c2) in G.SEQ_CASES.items():
        dv = d.put(np.full((12, V, 2), 7.0))
        d.c.qkxtm_seq_source(dv, ts, d1, d2 if part == 1 else None, 8, nu, c2, pid, particle, part)
        got = d.get(dv, (12, V, 2))
        assert np.all(got[:, :ts * V3] == 7.0) and np.all(got[:, (ts + 1) * V3:] == 7.0)        # only the time slice is written
        want = _c(gold[key])
        assert np.abs(_c(got[:, ts * V3:(ts + 1) * V3]) - want).max() / np.abs(want).max() < 1e-13, key
    for pid in range(5):
        for particle in range(2):
            for part, (nu, c2) in ((1, (1, 2)), (2, (2, 0))):
                dv = d.put(np.zeros((12, V, 2)))
                d.c.qkxtm_seq_source(dv, 0, d1, d2 if part == 1 else None, 8, nu, c2, pid, particle, part)
                got = _c(d.get(dv, (12, V, 2)))[:, :V3].reshape(4, 3, V3)
                want = O.seq_source_part1(_c(t1), _c(t2), nu, c2, pid, particle) if part == 1 else O.seq_source_part2(_c(t1), nu, c2, pid, particle)
                assert np.abs(got - want).max() / np.abs(want).max() < 1e-13, (pid, particle, part)
    # float 3-d propagators (what the driver uses)
    f1, f2 = d.put(t1.astype(np.float32)), d.put(t2.astype(np.float32))
    dv = d.put(np.zeros((12, V, 2), dtype=np.float32))
    d.c.qkxtm_seq_source(dv, ts, f1, f2, 4, 0, 0, 0, 0, 1)
    want = _c(gold["seq1_G4_proton_00"])
    assert np.abs(_c(d.get(dv, (12, V, 2), np.float32))[:, ts * V3:(ts + 1) * V3] - want).max() / np.abs(want).max() < 1e-5
    # ultra-local insertion
    moms = G.baryon_momenta()
    got = d.c.qkxtm_fixsink_local(d.put(p2), d.put(p1), 8, 0, 1, moms, G.SRC, X[3])            # (seq, fwd) = (p2, p1) as in the fixture call
    want = _c(gold["thrp_local_double"])
    assert _relmax(got, want) < 1e-13
    gotf = d.c.qkxtm_fixsink_local(d.put(p2.astype(np.float32)), d.put(p1.astype(np.float32)), 4, 1, 1, moms, G.SRC, X[3])
    assert _relmax(gotf, _c(gold["thrp_local_float"]).astype(np.complex128)) < 1e-5
    for particle in range(2):
        for pf in (1, 2):
            got = d.c.qkxtm_fixsink_local(d.bufs[-4], d.bufs[-3], 8, particle, pf, moms[:2], G.SRC, X[3])
            assert _relmax(got, O.fixsink_local_mom(_c(p1), _c(p2), X, moms[:2], G.SRC, particle, pf)) < 1e-13, (particle, pf)
    d.close()


def test_derivative_insertions_match_reference_fixture(tmq):
    """Noether and one-derivative insertions against the reference's kernel bodies (fixture) and, on a lattice with unequal
    extents and float inputs, against the restatement"""
    gold = np.load(G.FIXTURE)
    p1, p2 = G.contract_inputs()
    U = G.deriv_gauge()
    d = Dev(tmq, G.X)
    moms = G.baryon_momenta()
    gn, go = d.c.qkxtm_fixsink_derivative(d.put(p2), d.put(p1), d.put(U), 8, 1, 2, moms, G.SRC)       # (seq, fwd) = (p2, p1)
    assert _relmax(gn, _c(gold["thrp_noether_double"])) < 1e-13
    assert _relmax(go, _c(gold["thrp_oneD_double"])) < 1e-13
    d.c.set_option(tmq.OPT_CONTRACT_SLICES, 1)                        # one time slice per pass: the neighbours in t lie outside the pass
    gn1, go1 = d.c.qkxtm_fixsink_derivative(d.bufs[0], d.bufs[1], d.bufs[2], 8, 1, 2, moms, G.SRC)
    assert np.array_equal(gn1, gn) and np.array_equal(go1, go)
    d.close()
    X = (6, 4, 2, 8)
    rng = np.random.default_rng(31)
    V = int(np.prod(X))
    F = rng.standard_normal((4, 4, 3, 3, V, 2)).astype(np.float32); S = rng.standard_normal((4, 4, 3, 3, V, 2)).astype(np.float32)
    Ug = rng.standard_normal((4, 3, 3, V, 2)).astype(np.float32)
    d = Dev(tmq, X)
    mm, src = [(0, 0, 0), (1, -1, 0), (0, 2, 1)], (5, 1, 0)
    gn, go = d.c.qkxtm_fixsink_derivative(d.put(S), d.put(F), d.put(Ug), 4, 0, 1, mm, src)
    wn, wo = O.fixsink_derivative_mom(_c(F).astype(np.complex128), _c(S).astype(np.complex128), _c(Ug).astype(np.complex128), X, mm, src, 0, 1)
    assert _relmax(gn, wn) < 2e-5 and _relmax(go, wo) < 2e-5
    d.close()
    # a split lattice is refused (the neighbours' propagators would be needed)
    d = Dev(tmq, X)
    d.c.force_partition((0, 0, 0, 1))
    with pytest.raises(tmq.TmqError):
        d.c.qkxtm_fixsink_derivative(d.put(S), d.put(F), d.put(Ug), 4, 0, 1, mm, src)
    d.close()


def test_site_local_propagator_kernels_match_reference_fixture(tmq):
    gold = np.load(G.FIXTURE)
    p1, _ = G.contract_inputs()
    d = Dev(tmq, G.X)
    V = d.V
    for sign, key in ((+1, "rotate_plus"), (-1, "rotate_minus")):
        dp = d.put(p1)
        d.c.qkxtm_rotate_physical(dp, 8, sign)
        got = d.get(dp, p1.shape)
        assert np.abs(_c(got)[..., G.SAMPLE] - _c(gold[key])).max() < 1e-15
        assert np.abs(_c(got) - O.rotate_physical(_c(p1), sign)).max() < 1e-15
    df = d.put(p1.astype(np.float32))
    d.c.qkxtm_rotate_physical(df, 4, +1)
    assert np.abs(_c(d.get(df, p1.shape, np.float32))[..., G.SAMPLE] - _c(gold["rotate_plus_f32"])).max() < 1e-6
    dp = d.put(p1)
    d.c.qkxtm_gamma5_prop(dp, 8)
    got = d.get(dp, p1.shape)
    assert np.array_equal(got[..., G.SAMPLE, :], gold["gamma5_prop"]) and np.array_equal(got, p1[[2, 3, 0, 1]])
    d.c.qkxtm_gamma5_prop(dp, 8)
    d.c.qkxtm_conjugate(dp, 8, 144)
    got = d.get(dp, p1.shape)
    assert np.array_equal(got[..., G.SAMPLE, :], gold["conj_prop"])
    vec = np.ascontiguousarray(p1[:, 0, :, 0]).reshape(12, V, 2)
    dv = d.put(vec)
    d.c.qkxtm_conjugate(dv, 8, 12)
    assert np.array_equal(d.get(dv, vec.shape)[..., G.SAMPLE, :], gold["conj_vec"])
    with pytest.raises(tmq.TmqError):
        d.c.qkxtm_rotate_physical(dp, 8, 2)                     # "The sign can be only +-1" (lib/qudaQKXTM_Propagator.cpp:110)
    d.close()


def test_column_copies_between_propagator_and_vector(tmq):
    """absorbVectorToDevice / copyPropagator (whole volume) and copyPropagator3D / absorbVectorTimeSlice (one time slice),
    lib/qudaQKXTM_Vector.cpp:463-512, lib/qudaQKXTM_Propagator.cpp:90-106,533-550: bit-exact strided copies"""
    X = (4, 4, 2, 6)
    rng = np.random.default_rng(5)
    V = int(np.prod(X)); V3 = V // X[3]
    prop = rng.standard_normal((4, 4, 3, 3, V, 2)); vec = rng.standard_normal((4, 3, V, 2))
    d = Dev(tmq, X)
    dp, dv = d.put(prop), d.put(vec)
    nu, c2 = 2, 1
    d.c.qkxtm_column_copy(dp, V, 0, dv, V, 0, V, 8, nu, c2, True)
    want = prop.copy(); want[:, nu, :, c2] = vec
    assert np.array_equal(d.get(dp, prop.shape), want)
    assert np.array_equal(d.get(dp, prop.shape), _absorb_ref(d, tmq, prop, vec, nu, c2))
    dv2 = d.put(np.zeros_like(vec))
    d.c.qkxtm_column_copy(dp, V, 0, dv2, V, 0, V, 8, 3, 0, False)
    assert np.array_equal(d.get(dv2, vec.shape), want[:, 3, :, 0])
    # 3-d propagator <- one time slice of a vector, and back into another slice of a vector
    p3 = np.zeros((4, 4, 3, 3, V3, 2)); dp3 = d.put(p3)
    ts = 4
    d.c.qkxtm_column_copy(dp3, V3, 0, dv, V, ts * V3, V3, 8, nu, c2, True)
    got3 = d.get(dp3, p3.shape)
    assert np.array_equal(got3[:, nu, :, c2], vec[:, :, ts * V3:(ts + 1) * V3]) and np.count_nonzero(got3) == 12 * V3 * 2
    d.c.qkxtm_column_copy(dp3, V3, 0, dv2, V, 1 * V3, V3, 8, nu, c2, False)
    assert np.array_equal(d.get(dv2, vec.shape)[:, :, V3:2 * V3], vec[:, :, ts * V3:(ts + 1) * V3])
    with pytest.raises(tmq.TmqError):
        d.c.qkxtm_column_copy(dp3, V3, 0, dv, V, (X[3] - 1) * V3 + 1, V3, 8, nu, c2, True)      # runs off the vector
    d.close()


def _absorb_ref(d, tmq, prop, vec, nu, c2):
    dp, dv = d.put(prop), d.put(vec)
    d.c.qkxtm_absorb(dp, dv, 8, nu, c2)
    return d.get(dp, prop.shape)


# ---- the reference-shaped driver: calcMG_threepTwop_EvenOdd, meson two-point part -----------------------------------------------
XD = (4, 6, 4, 8)
KAPPA = 1.0 / (2.0 * 4.1)
MU = 0.1


def _oracle_propagator(o, gauge, src_lex, mu, smear=None):
    """12 columns of M_full(mu)^-1 on point sources at src_lex through the CPU oracle (prepare -> M^dag -> CG -> reconstruct);
    returns the propagator in the QKXTM layout [mu][nu][c1][c2][x_lex] complex"""
    V = o.V
    prop = np.zeros((4, 4, 3, 3, V), dtype=np.complex128)
    for isc in range(12):
        b = np.zeros((12, V), dtype=np.complex128); b[isc, src_lex] = 1.0
        if smear is not None:
            b = smear(b)
        b_lex = lu.c2r(np.transpose(b.reshape(4, 3, V), (2, 0, 1)))
        b_eo = lu.spinor_eo_from_lex(b_lex, XD)
        src = o.prepare(gauge, b_eo, KAPPA, mu)
        rhs = o.matpc(gauge, src, KAPPA, mu, 0, dagger=1)
        x_pc, _, _, _ = o.cg_mdagm(gauge, rhs, KAPPA, mu, tol=1e-13)
        x = np.zeros_like(b_eo); x[: V // 2] = x_pc
        o.reconstruct(gauge, x, b_eo, KAPPA, mu)
        col = lu.r2c(lu.spinor_lex_from_eo(x, XD))                                 # [x_lex][s][c]
        if smear is not None:
            col = np.transpose(smear(np.transpose(col, (1, 2, 0)).reshape(12, V)).reshape(4, 3, V), (2, 0, 1))
        prop[:, isc // 3, :, isc % 3] = np.transpose(col, (1, 2, 0))
    return prop


@pytest.mark.parametrize("nsmear", [0, 2])
def test_twop_driver_meson_correlators_match_oracle(tmp_path, tmq, nsmear):
    """qkxtm_invert_test --test twop = calcMG_threepTwop_EvenOdd (lib/qudaQKXTM_interface.cpp:236-1290, meson two-point part):
    24 solves (up / down), cast to float, [sink smearing,] rotateToPhysicalBase(+-1), contractMesons, ASCII file.  The oracle
    redoes it on the CPU: its own CG for the 24 columns, then the restatement pinned to the reference's contraction kernel."""
    from oracle.oracle import Oracle, gauss_smear
    o = Oracle(XD)
    gauge = tmq.gen_gauge(XD, seed=137, t_boundary=-1)
    src = (1, 3, 2, 5)
    q_sq, alpha = 2, 4.0
    out = str(tmp_path / "tw")
    cmd = [DRV, "--dim"] + [str(v) for v in XD] + ["--test", "twop", "--tol", "1e-11", "--recon", "12", "--Q_sq", str(q_sq), "--src"] + \
          [str(v) for v in src] + ["--nsmearGauss", str(nsmear), "--alphaGauss", str(alpha), "--out", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count(" up - ") == 12 and p.stdout.count(" dn - ") == 12
    fname = "%s.mesons.SS.%02d.%02d.%02d.%02d.dat" % ((out,) + src)
    rows = np.loadtxt(fname)
    moms = O.create_momenta(q_sq)
    T = XD[3]
    assert rows.shape == (10 * T * len(moms), 9)
    got = (rows[:, 5] + 1j * rows[:, 6]).reshape(10, T, len(moms)), (rows[:, 7] + 1j * rows[:, 8]).reshape(10, T, len(moms))
    assert np.array_equal(rows[: len(moms), 2:5], np.array(moms))

    V = int(np.prod(XD))
    smear = None
    if nsmear:
        U = lu.r2c(np.stack([lu.spinor_lex_from_eo(gauge[m], XD) for m in range(4)]))       # [4][x_lex][3][3], as the driver passes it
        Uq = np.transpose(U, (0, 2, 3, 1))
        smear = lambda v: gauss_smear(v, Uq, XD, alpha, nsmear)
    src_lex = ((src[3] * XD[2] + src[2]) * XD[1] + src[1]) * XD[0] + src[0]
    up = _oracle_propagator(o, gauge, src_lex, +MU, smear).astype(np.complex64).astype(np.complex128)     # K_temp is a float vector
    dn = _oracle_propagator(o, gauge, src_lex, -MU, smear).astype(np.complex64).astype(np.complex128)
    want = O.contract_mesons_mom(O.rotate_physical(up, +1), O.rotate_physical(dn, -1), XD, moms, src[:3])   # [T][nmoms][2][10]
    for iu in range(2):
        w = np.transpose(want[:, :, iu, :], (2, 0, 1))                               # [ip][t][imom]
        w = np.roll(w, -src[3], axis=1)                                              # time relative to the source
        assert _relmax(got[iu], w) < 2e-5, iu
    # the baryon file of the same run: "ip it px py pz gamma gammap re(iu=0) im(iu=0) re(iu=1) im(iu=1)", time relative to the source
    # with the sign flip of the anti-periodic boundary where it wraps (lib/qudaQKXTM_Contraction.cpp:886-898)
    from oracle import ref
    if nsmear == 0 and ref.available():
        rb = np.loadtxt("%s.baryons.SS.%02d.%02d.%02d.%02d.dat" % ((out,) + src))
        assert rb.shape == (10 * T * len(moms) * 16, 11)
        gb = [(rb[:, 7 + 2 * iu] + 1j * rb[:, 8 + 2 * iu]).reshape(10, T, len(moms), 4, 4) for iu in range(2)]
        # the reference's own kernel body (float, as the reference launches it) on the oracle's rotated float propagators
        c2r32 = lambda z: np.ascontiguousarray(np.stack([z.real, z.imag], axis=-1).astype(np.float32))
        bw = _c(ref.Ref(XD).contract_baryons_mom(c2r32(O.rotate_physical(up, +1)), c2r32(O.rotate_physical(dn, -1)), moms[:3], src[:3]))
        bw = bw.astype(np.complex128)                                               # [T][3 moms][2][10][4][4]
        for iu in range(2):
            w = np.transpose(bw[:, :, iu], (2, 0, 1, 3, 4))                       # [ip][t][imom][g][g']
            w = np.roll(w, -src[3], axis=1)
            w[:, T - src[3]:] *= -1.0
            for ip in range(10):
                assert _relmax(gb[iu][ip][:, :3], w[ip]) < 1e-4, (iu, ip)
    # physics: the pseudoscalar correlator at zero momentum is sum |S|^2: real and positive for both flavours, largest at the source
    for iu in range(2):
        pion = got[iu][0, :, 0]
        assert np.all(pion.real > 0) and np.abs(pion.imag).max() < 1e-6 * pion.real.max() and np.argmax(pion.real) == 0


@pytest.mark.parametrize("particle,proj", [("proton", 0), ("neutron", 3)])
def test_threep_driver_local_insertion_matches_oracle(tmp_path, tmq, particle, proj):
    """qkxtm_invert_test --test twop --tsink dt: the fixed-sink three-point function of calcMG_threepTwop_EvenOdd
    (lib/qudaQKXTM_interface.cpp:764-1170, ultra-local insertion): 24 forward solves, 3-d propagators at the sink, 12 + 12
    sequential sources -> conjugate, gamma5 -> solve with the other flavour -> K_seqProp -> contractFixSink -> ASCII files.
    The oracle redoes the whole chain on the CPU with its own CG and the restatements pinned to the reference's kernel bodies."""
    from oracle.oracle import Oracle
    o = Oracle(XD)
    gauge = tmq.gen_gauge(XD, seed=137, t_boundary=-1)
    src, dt, q_sq = (1, 3, 2, 5), 5, 1           # sink at t = (5 + 5) % 8 = 2: beyond the anti-periodic boundary -> sign flip in the file
    out = str(tmp_path / "tw")
    cmd = [DRV, "--dim"] + [str(v) for v in XD] + ["--test", "twop", "--tol", "1e-12", "--recon", "12", "--Q_sq", str(q_sq), "--src"] + \
          [str(v) for v in src] + ["--tsink", str(dt), "--proj", str(proj), "--particle", particle, "--out", out]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    T = XD[3]; V = int(np.prod(XD)); V3 = V // T
    moms = O.create_momenta(q_sq)
    part_id = 0 if particle == "proton" else 1
    pname = ["G4", "G5G123", "G5G1", "G5G2", "G5G3"][proj]
    src_lex = ((src[3] * XD[2] + src[2]) * XD[1] + src[1]) * XD[0] + src[0]
    f32 = lambda z: z.astype(np.complex64).astype(np.complex128)
    up, dn = f32(_oracle_propagator(o, gauge, src_lex, +MU)), f32(_oracle_propagator(o, gauge, src_lex, -MU))
    tsink = (src[3] + dt) % T
    up3, dn3 = up[..., tsink * V3:(tsink + 1) * V3], dn[..., tsink * V3:(tsink + 1) * V3]
    g5 = O._gamma5_ukqcd()
    Ulex = lu.r2c(np.stack([lu.spinor_lex_from_eo(gauge[m], XD) for m in range(4)]))      # [4][x_lex][3][3], what the driver passes as `gauge`
    Uq = np.ascontiguousarray(np.transpose(Ulex, (0, 2, 3, 1))).astype(np.complex64).astype(np.complex128)   # K_gaugeContractions is a float container
    for part in (1, 2):
        up_line = (particle == "proton") == (part == 1)
        mu_solve = -MU if up_line else +MU
        seq = np.zeros((4, 4, 3, 3, V), dtype=np.complex128)
        for nu in range(4):
            for c2 in range(3):
                if part == 1:
                    a, b = (up3, dn3) if particle == "proton" else (dn3, up3)
                    s3 = O.seq_source_part1(a, b, nu, c2, proj, part_id)
                else:
                    s3 = O.seq_source_part2(up3 if particle == "proton" else dn3, nu, c2, proj, part_id)
                s3 = f32(s3)                                              # K_temp is a float vector
                vec = np.zeros((4, 3, V), dtype=np.complex128); vec[:, :, tsink * V3:(tsink + 1) * V3] = np.einsum("ab,bcx->acx", g5, s3.conj())
                b_lex = lu.c2r(np.transpose(vec, (2, 0, 1)))
                b_eo = lu.spinor_eo_from_lex(np.ascontiguousarray(b_lex), XD)
                pc = o.prepare(gauge, b_eo, KAPPA, mu_solve)
                rhs = o.matpc(gauge, pc, KAPPA, mu_solve, 0, dagger=1)
                x_pc, _, _, _ = o.cg_mdagm(gauge, rhs, KAPPA, mu_solve, tol=1e-13)
                xf = np.zeros_like(b_eo); xf[: V // 2] = x_pc
                o.reconstruct(gauge, xf, b_eo, KAPPA, mu_solve)
                col = lu.r2c(lu.spinor_lex_from_eo(xf, XD))             # [x][s][c]
                seq[:, nu, :, c2] = np.transpose(col, (1, 2, 0))
        seq = f32(seq)
        fwd = up if up_line else dn
        want = O.fixsink_local_mom(fwd, seq, XD, moms, src[:3], part_id, part)        # [T][nmoms][16]
        flavour = "up" if up_line else "down"
        fname = "%s.threep_tsink%d_proj%s.%s.%s.ultra_local.SS.%02d.%02d.%02d.%02d.dat" % ((out, dt, pname, particle, flavour) + src)
        rows = np.loadtxt(fname)
        assert rows.shape == (16 * T * len(moms), 7)
        got = (rows[:, 5] + 1j * rows[:, 6]).reshape(16, T, len(moms))
        w = -np.roll(np.transpose(want, (2, 0, 1)), -src[3], axis=1)       # time relative to the source; src_t + dt >= T -> sign -1
        scale = np.abs(w).max()
        assert np.abs(got - w).max() / scale < 2e-4, (part, np.abs(got - w).max() / scale)
        # the conserved-current and one-derivative files of the same run (links: the configuration itself, as the driver passes it)
        wn, wo = O.fixsink_derivative_mom(fwd, seq, Uq.reshape(4, 3, 3, V), XD, moms, src[:3], part_id, part)
        rn = np.loadtxt(fname.replace("ultra_local", "noether")); ro = np.loadtxt(fname.replace("ultra_local", "oneD"))
        assert rn.shape == (4 * T * len(moms), 7) and ro.shape == (16 * 4 * T * len(moms), 8)
        gn = (rn[:, 5] + 1j * rn[:, 6]).reshape(4, T, len(moms))
        go = (ro[:, 6] + 1j * ro[:, 7]).reshape(16, 4, T, len(moms))
        wn_f = -np.roll(np.transpose(wn, (2, 0, 1)), -src[3], axis=1)                  # [dir][t][imom]
        wo_f = -np.roll(np.transpose(wo, (3, 2, 0, 1)), -src[3], axis=2)               # [iop][dir][t][imom]
        assert np.abs(gn - wn_f).max() / np.abs(wn_f).max() < 2e-4, part
        assert np.abs(go - wo_f).max() / np.abs(wo_f).max() < 2e-4, part
