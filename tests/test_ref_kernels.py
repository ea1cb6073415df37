"""CPU tests that pin the oracle's restatements of the plug-in's IN-TREE kernels against the reference itself:
the golden fixture tests/golden/qkxtm_ref_4x4x4x6.npz holds outputs of the reference's own kernel bodies
(lib/code_pieces/Gauss_core.h, uploadToCuda_core.h, downloadFromCuda_core.h, scaleVector_core.h,
apply_gamma5_vector_core.h) compiled for the CPU (oracle/ref_shim, tests/golden/make_golden_ref.py).  Where
oracle/_ref/libqkxtm_ref.so is present (this container; prebuilt on the GPU box) the library is also run live."""
import os
import sys

import numpy as np
import pytest

import lattice_util as lu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_ref as G  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(G.FIXTURE)


def _c(a):
    return a[..., 0] + 1j * a[..., 1]


def test_fixture_matches_live_reference_library(gold):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    vec, gauge = G.golden_inputs()
    r = ref.Ref(G.X, alpha_gauss=G.ALPHA)
    assert np.array_equal(r.gauss_step(vec, gauge), gold["gauss_step"])
    assert np.array_equal(r.gauss_smear(vec, gauge, G.NSMEAR), gold["gauss_smear3"])
    e, o = r.upload(vec)
    assert np.array_equal(e, gold["upload_even"]) and np.array_equal(o, gold["upload_odd"])


def test_gaussian_smearing_restatement_matches_reference(gold):
    from oracle.oracle import gauss_smear, gauss_smear_step
    vec, gauge = G.golden_inputs()
    one = gauss_smear_step(_c(vec), _c(gauge), G.X, G.ALPHA)
    assert lu.rel_l2(one, _c(gold["gauss_step"])) < 1e-15
    three = gauss_smear(_c(vec), _c(gauge), G.X, G.ALPHA, G.NSMEAR)
    assert lu.rel_l2(three, _c(gold["gauss_smear3"])) < 1e-15
    assert lu.rel_l2(one, _c(gold["gauss_step_f32"]).astype(np.complex128)) < 1e-6
    # nsmear = 0 copies (the cudaMemcpy of the even-count branch, lib/qudaQKXTM_Vector.cpp:419)
    assert np.array_equal(gauss_smear(_c(vec), _c(gauge), G.X, G.ALPHA, 0), _c(vec))


def test_upload_download_layout_restatement_matches_reference(gold):
    """QKXTM device layout [(s*3+c)][x_lex] <-> QUDA native order [(s*3+c)][cb] per parity:
    lattice_util's even-odd permutation, which the GPU converters are tested against, reproduces the reference kernels"""
    vec, _ = G.golden_inputs()
    V = int(np.prod(G.X)); Vh = V // 2
    lex = np.ascontiguousarray(np.transpose(vec.reshape(4, 3, V, 2), (2, 0, 1, 3)))      # [x][s][c][ri]
    eo = lu.spinor_eo_from_lex(lex, G.X)                                                  # [even Vh | odd Vh][s][c][ri]
    even = np.transpose(eo[:Vh], (1, 2, 0, 3)).reshape(12, Vh, 2)
    odd = np.transpose(eo[Vh:], (1, 2, 0, 3)).reshape(12, Vh, 2)
    assert np.array_equal(even, gold["upload_even"]) and np.array_equal(odd, gold["upload_odd"])
    # download: inverse; an absent parity is zero-filled
    both = gold["download_both"]
    assert np.array_equal(both, vec)
    only = gold["download_even_only"].reshape(4, 3, V, 2)
    lex_only = np.transpose(only, (2, 0, 1, 3))
    eo_only = lu.spinor_eo_from_lex(np.ascontiguousarray(lex_only), G.X)
    assert np.array_equal(eo_only[:Vh], eo[:Vh]) and np.all(eo_only[Vh:] == 0)


def test_scale_and_gamma5_restatement_matches_reference(gold):
    from oracle.oracle import Oracle
    vec, _ = G.golden_inputs()
    assert np.array_equal(gold["scale"], (2 * 0.1234) * vec)
    # gamma5 in the UKQCD basis swaps spins 0<->2, 1<->3 (apply_gamma5_vector_core.h) = the oracle's g1 g2 g3 g4
    V = int(np.prod(G.X))
    g5 = Oracle(G.X).gamma5()
    v = _c(vec).reshape(4, 3, V)
    assert np.allclose(np.einsum("st,tcx->scx", g5, v), _c(gold["gamma5"]).reshape(4, 3, V), atol=0, rtol=0)


def test_plaquette_restatement_matches_reference(gold):
    """QKXTM_Gauge::calculatePlaq: the reference's plaquette kernel (block reduction emulated with one OS thread per CUDA
    thread) against the oracle's plaquette on the same links in QDP even-odd order; the fixture uses the golden gauge field
    (general complex matrices, not SU(3): the kernel is a polynomial in the links)"""
    from oracle.oracle import Oracle
    _, gauge = G.golden_inputs()
    V = int(np.prod(G.X))
    U_lex = np.transpose(_c(gauge), (0, 3, 1, 2))                    # [4][V][3][3]
    gq = lu.gauge_qdp_from_lex(U_lex, G.X, t_boundary=+1)
    assert abs(Oracle(G.X).plaquette(gq) - float(gold["plaquette"][0])) < 1e-13 * abs(float(gold["plaquette"][0]))
    from oracle import ref
    if ref.available():
        assert ref.Ref(G.X).plaquette(gauge) == float(gold["plaquette"][0])
