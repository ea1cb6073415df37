"""CPU tests that pin the geometry / gauge conventions of the oracle and of the host layer
against the REFERENCE'S OWN host code: qkxtm/QKXTM_util.cpp and include/QKXTM_read_conf.h compiled in place into
oracle/_ref/libqkxtm_util_ref.so (oracle/ref_shim/qkxtm_util_host.cpp; upstream-QUDA headers replaced by declaration-only
stand-ins).  Skipped where the library has not been built (it needs /root/reference at build time; it travels prebuilt)."""
import numpy as np
import pytest

import lattice_util as lu
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.util_available(), reason="oracle/_ref/libqkxtm_util_ref.so not built")

X = (4, 6, 2, 8)
DIRS = [(0, 0, 0, 1), (0, 0, 0, -1), (0, 0, 1, 0), (0, 0, -1, 0), (0, 1, 0, 0), (0, -1, 0, 0), (1, 0, 0, 0), (-1, 0, 0, 0)]


def test_index_helpers_match_the_reference():
    from oracle.oracle import Oracle
    r = ref.RefUtil(X); o = Oracle(X)
    perm = lu.eo_from_lex(X)                     # lexicographic index of every even-odd-ordered site
    for odd in (0, 1):
        for i in range(r.Vh):
            Y = r.full_index(i, odd)
            assert Y == o.full_index(i, odd) == perm[odd * r.Vh + i]
            assert r.odd_bit(Y) == odd
            for d in DIRS:
                assert r.neighbor_index(i, odd, *d) == o.neighbor_index(i, odd, *d)


def test_reconstruct12_and_boundary_sign_match_the_reference():
    from oracle.oracle import Oracle
    r = ref.RefUtil(X); o = Oracle(X)
    rng = np.random.default_rng(2)
    last = (X[3] - 1) * X[0] * X[1] * X[2] // 2          # first checkerboard index of the last time slice
    for direction, ga_idx, tb in [(0, 5, -1), (2, last + 3, -1), (3, 5, -1), (3, last, -1), (3, r.Vh - 1, -1), (3, last + 1, 1)]:
        m = rng.standard_normal(18)
        u0 = 1.0 if direction < 3 else (float(tb) if ga_idx >= last else 1.0)
        assert np.array_equal(r.reconstruct12(m, direction, ga_idx, tb), o.reconstruct12(m, u0))


def test_reference_random_field_is_su3_and_recon12_restores_it():
    """constructGaugeField (qkxtm/QKXTM_util.cpp:879-955) run from the reference: every link is special unitary, the anti-
    periodic sign sits on U_t of the last time slice, and rebuilding the third row from the first two with the reference's
    su3Reconstruct12 convention -- the rule the device applies in registers -- returns the stored third row"""
    from oracle.oracle import Oracle
    r = ref.RefUtil(X); o = Oracle(X)
    g = r.construct_gauge_field(1, seed=137, t_boundary=-1)
    U = lu.r2c(g)
    last = (X[3] - 1) * X[0] * X[1] * X[2] // 2
    for mu in range(4):
        UU = np.einsum("xab,xcb->xac", U[mu], np.conj(U[mu]))
        assert np.abs(UU - np.eye(3)).max() < 1e-13
        det = np.linalg.det(U[mu])
        sign = np.ones(r.V)
        if mu == 3:
            for par in (0, 1):
                sign[par * r.Vh + last: (par + 1) * r.Vh] = -1.0      # det(-U) = -1 for a 3x3 matrix
        assert np.abs(det - sign).max() < 1e-13
        for site in (0, 7, r.Vh + last + 2, r.V - 1):
            m = g[mu, site].reshape(18).copy()
            want = m[12:].copy()
            u0 = -1.0 if (mu == 3 and (site % r.Vh) >= last) else 1.0
            assert np.abs(o.reconstruct12(m, u0)[12:] - want).max() < 1e-14


def test_boundary_condition_matches_the_reference():
    import tmq
    r = ref.RefUtil(X)
    U = lu.random_su3_lex(X, seed=9)
    periodic = lu.gauge_qdp_from_lex(U, X, t_boundary=+1)
    want = r.apply_gauge_field_scaling(periodic, -1)                     # applyGaugeFieldScaling, anisotropy 1
    assert np.array_equal(want, lu.gauge_qdp_from_lex(U, X, t_boundary=-1))
    mine = periodic.copy()
    tmq.apply_t_boundary(mine, X, t_boundary=-1)                          # the host library's applyBoundaryCondition
    assert np.array_equal(mine, want)
    assert np.array_equal(r.apply_gauge_field_scaling(periodic, +1), periodic)


def test_lime_reader_matches_the_reference_reader(tmp_path):
    """the reference's readLimeGauge (include/QKXTM_read_conf.h:107-400, single-rank branch) and host/tmq_lime.cpp read the
    same configuration file -- one written by tmq_lime_write_gauge, one assembled byte by byte in numpy -- identically"""
    import tmq
    from test_lime_io import build_ildg
    r = ref.RefUtil(X)
    U = lu.random_su3_lex(X, seed=4)
    g = lu.gauge_qdp_from_lex(U, X, t_boundary=+1)
    a = str(tmp_path / "written_by_tmq.lime")
    tmq.lime_write_gauge(a, g, X, kappa=0.1373, mu=0.004)
    b = str(tmp_path / "built_in_numpy.lime")
    build_ildg(b, U, X, 0.1373, 0.004)
    for path in (a, b):
        got_ref, Xr = r.read_lime_gauge(path)
        assert Xr == X
        assert np.array_equal(got_ref, g)
        assert np.array_equal(tmq.lime_read_gauge(path, X), got_ref)
