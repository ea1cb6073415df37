"""CPU tests of the host-side product code (libqkxtm_tmq.so field generators) and of the sharding helpers."""
import numpy as np
import pytest

import lattice_util as lu


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    T.load_host()
    return T


@pytest.mark.parametrize("grid,coord", [((1, 1, 1, 1), (0, 0, 0, 0)), ((1, 1, 2, 2), (0, 0, 1, 0)), ((1, 1, 1, 4), (0, 0, 0, 3))])
def test_cxx_generators_match_numpy(tmq, grid, coord):
    X = (4, 6, 4, 8)
    a = tmq.gen_gauge(X, grid=grid, coord=coord); b = lu.random_gauge_qdp(X, grid=grid, coord=coord)
    assert np.abs(a - b).max() < 1e-14
    a = tmq.gen_spinor(X, grid=grid, coord=coord); b = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, grid=grid, coord=coord), X)
    assert np.abs(a - b).max() < 1e-14
    a = tmq.gen_spinor(X, "z4", grid=grid, coord=coord, eo_order=False)
    assert np.array_equal(a, lu.z4_source_lex(X, grid=grid, coord=coord))
    assert set(np.unique(a)) <= {-1.0, 0.0, 1.0} and np.all(np.sum(np.abs(a), axis=-1) == 1.0)


def test_generated_links_are_su3_and_boundary_is_folded(tmq):
    X = (4, 4, 4, 6)
    g = lu.r2c(tmq.gen_gauge(X, t_boundary=-1))
    gp = lu.r2c(tmq.gen_gauge(X, t_boundary=+1))
    UUd = np.einsum("mxab,mxcb->mxac", g, np.conj(g))
    assert np.abs(UUd - np.eye(3)).max() < 1e-14
    assert np.abs(np.linalg.det(gp) - 1.0).max() < 1e-13
    # anti-periodic T: only U_t on the last time slice flips sign (qkxtm/QKXTM_util.cpp:698-705)
    x, y, z, t = lu.coords_lex(X)
    last = (t == X[3] - 1)[lu.eo_from_lex(X)]
    assert np.array_equal(g[:3], gp[:3])
    assert np.array_equal(g[3][last], -gp[3][last]) and np.array_equal(g[3][~last], gp[3][~last])
    unit = tmq.gen_gauge(X, unit=True, t_boundary=+1)
    assert np.array_equal(lu.r2c(unit)[2, 7], np.eye(3))


@pytest.mark.parametrize("grid", [(1, 1, 1, 2), (1, 1, 2, 1), (1, 1, 2, 4)])
def test_sharded_fields_are_slabs_of_the_global_field(tmq, grid):
    X = (4, 4, 2, 4)
    G = tuple(X[d] * grid[d] for d in range(4))
    psi_g = tmq.gen_spinor(G)
    gauge_g = tmq.gen_gauge(G)
    seen = np.zeros(int(np.prod(G)), dtype=np.int64)
    for rank in range(int(np.prod(grid))):
        coord = lu.rank_coord(rank, grid)
        assert lu.coord_rank(coord, grid) == rank
        assert np.array_equal(lu.local_from_global_eo(psi_g, X, grid, coord), tmq.gen_spinor(X, grid=grid, coord=coord))
        loc = tmq.gen_gauge(X, grid=grid, coord=coord)
        for mu in range(4):
            assert np.array_equal(lu.local_from_global_eo(gauge_g[mu], X, grid, coord), loc[mu])
        gl, _ = lu.global_lex(X, grid, coord)
        seen[gl.astype(np.int64)] += 1
    assert np.all(seen == 1)                      # the shards tile the global lattice exactly once


def test_bench_byte_model_matches_the_survey_table():
    import bench
    assert bench.bytes_per_site(0, 8, 12) == 1152 and bench.bytes_per_site(0, 8, 18) == 1536
    assert bench.bytes_per_site(2, 8, 12) == 1344 and bench.bytes_per_site(2, 4, 12) == 672
    assert bench.step_bytes_per_site(8, 12) == (16 * 24 + 4 * 96) * 8
    assert bench.choose_grid(8) == (1, 1, 1, 8)
