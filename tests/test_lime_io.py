"""CPU tests of the on-disk formats (host/tmq_lime.cpp): the ILDG / LIME configuration reader of
include/QKXTM_read_conf.h:107-400 and the DiracFermion_Sink writer of lib/qudaQKXTM_Vector.cpp:510-702, checked against
the format itself -- a file assembled byte by byte in numpy from the record layout (144-byte big-endian LIME header,
payload padded to 8 bytes; ildg-binary-data = big-endian doubles [t][z][y][x][mu][3][3][2]) -- and by round trips,
including reading one file as a 2 x 2 (z, t) process grid."""
import os
import struct

import numpy as np
import pytest

import lattice_util as lu


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    T.load_host()
    return T


def lime_record(rtype, payload, mb, me):
    h = struct.pack(">IHBBQ", 0x456789AB, 1, (0x80 if mb else 0) | (0x40 if me else 0), 0, len(payload))
    h += rtype.encode().ljust(128, b"\0")
    assert len(h) == 144
    return h + payload + b"\0" * ((8 - len(payload) % 8) % 8)


def build_ildg(path, U_lex, X, kappa, mu):
    """U_lex: complex [4][V][3][3], lexicographic x fastest"""
    V = int(np.prod(X))
    data = np.empty((V, 4, 3, 3, 2), dtype=">f8")
    data[..., 0] = np.transpose(U_lex, (1, 0, 2, 3)).real
    data[..., 1] = np.transpose(U_lex, (1, 0, 2, 3)).imag
    xlf = ("plaquette = 0.5\n trajectory nr = 7\n beta = 1.95, kappa = %.10f, mu = %.10f, c2_rec = -0.083\n" % (kappa, mu)).encode()
    fmt = ("<?xml version=\"1.0\" encoding=\"UTF-8\"?><ildgFormat><version>1.0</version><field>su3gauge</field><precision>64</precision>"
           "<lx>%d</lx><ly>%d</ly><lz>%d</lz><lt>%d</lt></ildgFormat>" % X).encode()
    with open(path, "wb") as f:
        f.write(lime_record("xlf-info", xlf, True, True))
        f.write(lime_record("ildg-format", fmt, True, False))
        f.write(lime_record("ildg-binary-data", data.tobytes(), False, False))
        f.write(lime_record("ildg-data-lfn", b"lfn://synthetic", False, True))


def test_read_ildg_configuration_built_from_the_format(tmq, tmp_path):
    X = (4, 6, 4, 8)
    U = lu.random_su3_lex(X, seed=21)
    path = str(tmp_path / "conf.0000")
    build_ildg(path, U, X, 0.1373, 0.004)
    info = tmq.lime_gauge_info(path)
    assert info["X"] == X and info["precision"] == 64
    assert abs(info["kappa"] - 0.1373) < 1e-12 and abs(info["mu"] - 0.004) < 1e-12
    g = tmq.lime_read_gauge(path, X)
    want = lu.gauge_qdp_from_lex(U, X, t_boundary=+1)                 # the reader applies no boundary condition
    assert np.array_equal(g, want)
    # applyBoundaryCondition afterwards = the anti-periodic field the solver takes
    tmq.apply_t_boundary(g, X, t_boundary=-1)
    assert np.array_equal(g, lu.gauge_qdp_from_lex(U, X, t_boundary=-1))
    assert np.array_equal(g, lu.random_gauge_qdp(X, seed=21, t_boundary=-1))


def test_read_as_process_grid_and_write_round_trip(tmq, tmp_path):
    GX, grid = (4, 4, 8, 8), (1, 1, 2, 2)
    X = tuple(GX[d] // grid[d] for d in range(4))
    U = lu.random_su3_lex(GX, seed=5)
    path = str(tmp_path / "conf.big")
    build_ildg(path, U, GX, 0.125, 0.1)
    whole = tmq.lime_read_gauge(path, GX)
    out = str(tmp_path / "conf.rewritten")
    for r in range(4):
        coord = lu.rank_coord(r, grid)
        loc = tmq.lime_read_gauge(path, X, grid, coord)
        assert np.array_equal(loc, lu.random_gauge_qdp(X, seed=5, t_boundary=+1, grid=grid, coord=coord))
        tmq.lime_write_gauge(out, loc, X, grid, coord, kappa=0.125, mu=0.1)        # rank 0 first: it creates the file
    assert np.array_equal(tmq.lime_read_gauge(out, GX), whole)
    assert tmq.lime_gauge_info(out)["X"] == GX


def test_diracfermion_sink_writer(tmq, tmp_path):
    X = (4, 4, 6, 4)
    V = int(np.prod(X))
    psi = lu.gaussian_spinor_lex(X, seed=3)                            # [x_lex][4][3][2]
    for dt in (np.float64, np.float32):
        path = str(tmp_path / ("prop_%d.lime" % np.dtype(dt).itemsize))
        tmq.lime_write_vector(path, psi.astype(dt), X)
        raw = open(path, "rb").read()
        # three records: propagator-type, quda-propagator-format, scidac-binary-data (lib/qudaQKXTM_Vector.cpp:545-628)
        pos, types = 0, []
        while pos < len(raw):
            magic, ver, flags, _, n = struct.unpack(">IHBBQ", raw[pos:pos + 16])
            assert magic == 0x456789AB and ver == 1
            types.append((raw[pos + 16:pos + 144].rstrip(b"\0").decode(), pos + 144, n))
            pos += 144 + ((n + 7) // 8) * 8
        assert [t[0] for t in types] == ["propagator-type", "quda-propagator-format", "scidac-binary-data"]
        assert raw[types[0][1]:types[0][1] + types[0][2]] == b"DiracFermion_Sink"
        fmt = raw[types[1][1]:types[1][1] + types[1][2]].decode()
        assert "<precision>%d</precision>" % (8 * np.dtype(dt).itemsize) in fmt and "<lx>4</lx>" in fmt and "<lt>4</lt>" in fmt
        be = np.frombuffer(raw[types[2][1]:types[2][1] + types[2][2]], dtype=np.dtype(dt).newbyteorder(">")).reshape(V, 4, 3, 2)
        assert np.array_equal(be.astype(dt), psi.astype(dt))
        assert np.array_equal(tmq.lime_read_vector(path, X, dt), psi.astype(dt))


def test_errors(tmq, tmp_path):
    with pytest.raises(tmq.TmqError):
        tmq.lime_gauge_info(str(tmp_path / "missing"))
    bad = str(tmp_path / "bad")
    open(bad, "wb").write(b"\0" * 200)
    with pytest.raises(tmq.TmqError):
        tmq.lime_gauge_info(bad)
    X = (4, 4, 4, 4)
    path = str(tmp_path / "c")
    build_ildg(path, lu.random_su3_lex(X, seed=1), X, 0.1, 0.0)
    with pytest.raises(tmq.TmqError):
        tmq.lime_read_gauge(path, (4, 4, 4, 8))                       # wrong extents


def _write_worker(rank, world, port, path, X, grid, q):
    """one rank of QKXTM_Vector::write on a process grid (host/qudaQKXTM_tmq.cpp): rank 0 creates the file, ALL ranks meet at a barrier,
    every rank writes its block, barrier -- the order the reference gets from its MPI_Bcast of the payload offset (lib/qudaQKXTM_Vector.cpp:625)"""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tmq as T
        coord = lu.rank_coord(rank, grid)
        psi = lu.gaussian_spinor_lex(X, seed=3, grid=grid, coord=coord)
        if rank == 0:
            T.lime_write_vector_header(path, X, grid)
        dist.barrier()
        T.lime_write_vector_block(path, psi, X, grid, coord)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(1, 1, 1, 2), (1, 1, 2, 1)])
def test_two_ranks_write_a_propagator_over_a_stale_file(tmq, tmp_path, grid):
    """world_size-2 (gloo) run of the two-step writer over an EXISTING file of the same name with other content and size: the result must be
    the complete new propagator (without the barrier a non-root rank could write into the old file, which rank 0 then truncates)"""
    import socket
    import torch.multiprocessing as mp
    X = (4, 4, 4, 4)
    G = tuple(X[d] * grid[d] for d in range(4))
    path = str(tmp_path / "prop.lime")
    tmq.lime_write_vector(path, lu.gaussian_spinor_lex((4, 4, 4, 6), seed=9), (4, 4, 4, 6))     # the stale file: another lattice
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_write_worker, args=(r, 2, port, path, X, grid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
    want = lu.gaussian_spinor_lex(G, seed=3)
    assert np.array_equal(tmq.lime_read_vector(path, G), want)
    # a block written against a file whose payload does not match the lattice is refused, not silently accepted
    tmq.lime_write_vector(path, lu.gaussian_spinor_lex((4, 4, 4, 6), seed=9), (4, 4, 4, 6))
    with pytest.raises(tmq.TmqError, match="stale"):
        tmq.lime_write_vector_block(path, want[: int(np.prod(X))], X, grid, (0, 0, 0, 0))
