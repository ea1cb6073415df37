"""CPU tests that pin the oracle's restatement of the step AFTER the solves -- the propagator container kernels and the meson
two-point contraction of calcMG_threepTwop_EvenOdd (lib/qudaQKXTM_interface.cpp:1190-1223) -- against the reference itself:
tests/golden/qkxtm_ref_contract_4x4x4x6.npz holds outputs of the reference's own kernel bodies (contractMesons_core.h with
the tables of lib/qudaQKXTM_kernels.cu:77-78, rotateToPhysicalBase_core.h, apply_gamma5_propagator_core.h, conjugate_*_core.h)
compiled for the CPU (oracle/ref_shim, tests/golden/make_golden_contract.py).  Where oracle/_ref/libqkxtm_ref.so is present
the library is also run live."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_contract as G  # noqa: E402

from oracle import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(G.FIXTURE)


def _c(a):
    return a[..., 0] + 1j * a[..., 1]


def test_fixture_matches_live_reference_library(gold):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    p1, p2 = G.contract_inputs()
    r = ref.Ref(G.X)
    assert np.array_equal(r.contract_mesons_mom(p1, p2, G.momenta(), G.SRC), gold["mom_double"])
    assert np.array_equal(r.contract_mesons_mom(p1.astype(np.float32), p2.astype(np.float32), G.momenta(), G.SRC), gold["mom_float"])
    assert np.array_equal(r.rotate_physical(p1, -1)[..., G.SAMPLE, :], gold["rotate_minus"])


def test_meson_contraction_restatement_matches_reference(gold):
    """tr[G S G^dag g5 S^dag g5] with the reference's channel order and signs + the Fourier sum reproduce the reference's
    table-driven kernel: 1e-14 against its double instantiation, float rounding against the float one it launches"""
    p1, p2 = G.contract_inputs()
    want = O.contract_mesons_mom(_c(p1), _c(p2), G.X, G.momenta(), G.SRC)
    scale = np.abs(want).max()
    assert np.abs(_c(gold["mom_double"]) - want).max() / scale < 1e-14
    assert np.abs(_c(gold["mom_float"]).astype(np.complex128) - want).max() / scale < 2e-6
    site = np.stack([O.contract_mesons_site(_c(p1)), O.contract_mesons_site(_c(p2))])           # [2][10][V]
    pos = np.transpose(site.reshape(2, 10, G.X[3], -1), (2, 3, 0, 1))                            # [T][V3][2][10]
    assert np.abs(_c(gold["pos_float"]).astype(np.complex128) - pos).max() / np.abs(pos).max() < 2e-6


def test_meson_channels_are_what_their_names_say():
    """pseudoscalar = sum |S|^2 (real, positive); zero momentum of the projection = plain sum; the ten channels are real
    combinations for a g5-hermitian propagator"""
    p1, _ = G.contract_inputs()
    S = _c(p1)
    site = O.contract_mesons_site(S)
    assert np.allclose(site[0], (np.abs(S) ** 2).sum(axis=(0, 1, 2, 3)), rtol=1e-13)
    mom = O.contract_mesons_mom(S, S, G.X, [(0, 0, 0)], (0, 0, 0))
    assert np.allclose(mom[:, 0, 0, :], site.reshape(10, G.X[3], -1).sum(axis=2).T, rtol=1e-12)
    assert np.array_equal(mom[:, :, 0], mom[:, :, 1])
    # a shift of the source position is a pure phase per momentum
    m = [(1, 0, -1)]
    a = O.contract_mesons_mom(S, S, G.X, m, (0, 0, 0)); b = O.contract_mesons_mom(S, S, G.X, m, (1, 2, 3))
    ph = np.exp(2j * np.pi * (1 * 1 / G.X[0] + 0 - 1 * 3 / G.X[2]))
    assert np.allclose(b, a * ph, rtol=1e-12, atol=1e-12)


def test_momentum_list_order():
    """createMomenta (lib/qudaQKXTM_kernels.cu:98-116): shells of p^2 = 0, 1, 2, ... each scanned from +iQ down to -iQ"""
    assert [len(O.create_momenta(q)) for q in range(5)] == [1, 7, 19, 27, 33]
    m = O.create_momenta(1)
    assert m[0] == (0, 0, 0) and m[1] == (1, 0, 0) and m[-1] == (-1, 0, 0)


def test_site_local_propagator_kernels_match_reference(gold):
    p1, _ = G.contract_inputs()
    P = _c(p1)
    for sign, key in ((+1, "rotate_plus"), (-1, "rotate_minus")):
        got = O.rotate_physical(P, sign)[..., G.SAMPLE]
        assert np.abs(got - _c(gold[key])).max() < 1e-15
    assert np.abs(O.rotate_physical(P.astype(np.complex64), +1)[..., G.SAMPLE] - _c(gold["rotate_plus_f32"])).max() < 1e-6
    assert np.array_equal(P[[2, 3, 0, 1]][..., G.SAMPLE], _c(gold["gamma5_prop"]))             # gamma5 = spin swap on the sink index
    assert np.array_equal(P.conj()[..., G.SAMPLE], _c(gold["conj_prop"]))
    assert np.array_equal(P[:, 0, :, 0].reshape(12, -1).conj()[..., G.SAMPLE], _c(gold["conj_vec"]))
    # the rotation is an involution up to the twist: rotating with +1 then -1 gives back 1/4 (1 + g5 g5) P (1 + ...) = P
    back = O.rotate_physical(O.rotate_physical(P, +1), -1)
    assert np.abs(back - P).max() < 1e-14


def test_baryon_contraction_restatement_matches_reference(gold):
    """the ten baryon channels written as Gs x conj(Gr) x Xs x Xr gamma structures with explicit Wick terms (oracle.baryon_channels)
    reproduce the reference's table-driven kernel body (its double instantiation) to rounding, on a tiny lattice (the einsum
    restatement is slow); the big fixture entries are what the GPU kernel is compared with"""
    s1, s2 = G.small_inputs()
    want = O.contract_baryons_mom(_c(s1), _c(s2), G.X_SMALL, [(0, 0, 0), (1, 0, -1)], G.SRC_SMALL)
    got = _c(gold["baryon_small_double"])
    assert got.shape == want.shape == (2, 2, 2, 10, 4, 4)
    for ip in range(10):
        for iu in range(2):
            assert np.abs(got[:, :, iu, ip] - want[:, :, iu, ip]).max() / np.abs(want[:, :, iu, ip]).max() < 1e-13, (ip, iu)
    # float (what the reference launches) vs double instantiation of the same body on the larger fixture lattice
    d, f = _c(gold["baryon_mom_double"]), _c(gold["baryon_mom_float"]).astype(np.complex128)
    assert np.abs(d - f).max() / np.abs(d).max() < 2e-5


def test_baryon_fixture_matches_live_reference_library(gold):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    s1, s2 = G.small_inputs()
    assert np.array_equal(ref.Ref(G.X_SMALL).contract_baryons_mom(s1, s2, [(0, 0, 0), (1, 0, -1)], G.SRC_SMALL), gold["baryon_small_double"])


def test_threep_tables_are_twisted_rotations_of_the_physical_structures(gold):
    """projectors_tm_base.h = 1/4 (1 + g4)[i g5 g_k] and gammas_tm_base.h = the 16 insertions, both rotated with 1/2 (1 + i s g5) . (1 + i s g5)"""
    for pid in range(5):
        for part in range(2):
            assert np.abs(O.projector_tm(pid, part) - gold["proj_tables"][pid, part]).max() < 1e-15, (pid, part)
    for f in range(16):
        for part in range(2):
            for ipf, pf in enumerate((1, 2)):
                assert np.abs(O.operator_tm(f, part, pf) - gold["op_tables"][f, part, ipf]).max() < 1e-15, (f, part, pf)


def test_sequential_sources_and_local_insertion_match_reference(gold):
    p1, p2 = G.contract_inputs()
    V3 = int(np.prod(G.X[:3]))
    t1, t2 = _c(p1)[..., 2 * V3:3 * V3], _c(p2)[..., 2 * V3:3 * V3]
    for key, (part, pid, particle, nu, c2) in G.SEQ_CASES.items():
        want = _c(gold[key]).reshape(4, 3, V3)
        got = O.seq_source_part1(t1, t2, nu, c2, pid, particle) if part == 1 else O.seq_source_part2(t1, nu, c2, pid, particle)
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-13, key
    want = _c(gold["thrp_local_double"])
    got = O.fixsink_local_mom(_c(p1), _c(p2), G.X, G.baryon_momenta(), G.SRC, 0, 1)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-13
    wantf = _c(gold["thrp_local_float"]).astype(np.complex128)
    gotf = O.fixsink_local_mom(_c(p1), _c(p2), G.X, G.baryon_momenta(), G.SRC, 1, 1)
    assert np.abs(gotf - wantf).max() / np.abs(wantf).max() < 2e-6


def test_derivative_insertions_match_reference(gold):
    """conserved-current (Noether) and one-derivative insertions: four hop blocks per direction with the reference's 1/4"""
    p1, p2 = G.contract_inputs()
    wn, wo = O.fixsink_derivative_mom(_c(p1), _c(p2), _c(G.deriv_gauge()), G.X, G.baryon_momenta(), G.SRC, 1, 2)
    gn, go = _c(gold["thrp_noether_double"]), _c(gold["thrp_oneD_double"])
    assert np.abs(gn - wn).max() / np.abs(wn).max() < 1e-13
    assert np.abs(go - wo).max() / np.abs(wo).max() < 1e-13
