"""world_size-2 gloo test (CPU) of the sharded Dslash protocol libtmq implements on the device (SURVEY.md 8e,
csrc/tmq_halo.cu): each rank owns a slab, packs spin-projected half-spinor faces -- the forward-going face
already multiplied by U^dagger on the sender -- exchanges them with its neighbours, and adds the ghost terms on
its boundary slices.  The numpy restatement below runs that protocol over torch.distributed/gloo and must
reproduce the slab of the global dense operator, including the anti-periodic boundary link on the last rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lattice_util as lu
from oracle.oracle import gamma_ukqcd


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _lex_to_grid(a, X):
    return a.reshape((X[3], X[2], X[1], X[0]) + a.shape[1:])


def sharded_hop(U, psi, X, grid, coord, rank, gam, dagger):
    """U: [4][Vloc][3][3] local links (boundary sign folded in), psi: [Vloc][4][3]; returns D psi on the slab"""
    s = -1.0 if dagger else 1.0
    one = np.eye(4)
    P = _lex_to_grid(psi, X)
    out = np.zeros_like(P)
    for mu in range(4):
        ax = 3 - mu
        Um = _lex_to_grid(U[mu], X)
        Pm, Pp = one - s * gam[mu], one + s * gam[mu]
        fwd = np.einsum("st,...tc->...sc", Pm, P)                               # (1 - s g) psi
        bwd = np.einsum("...ba,...sb->...sa", np.conj(Um), np.einsum("st,...tc->...sc", Pp, P))   # U^dag (1 + s g) psi
        f_sh, b_sh = np.roll(fwd, -1, axis=ax), np.roll(bwd, +1, axis=ax)
        if grid[mu] > 1:
            lo = [slice(None)] * 4; hi = [slice(None)] * 4
            lo[ax] = slice(0, 1); hi[ax] = slice(X[mu] - 1, X[mu])
            send_bwd = torch.from_numpy(np.ascontiguousarray(fwd[tuple(lo)]))   # slice 0 -> rank-1 (its forward hop)
            send_fwd = torch.from_numpy(np.ascontiguousarray(bwd[tuple(hi)]))   # slice L-1 -> rank+1 (its backward hop)
            cm = list(coord); cm[mu] = (coord[mu] - 1) % grid[mu]; rm = lu.coord_rank(cm, grid)
            cp = list(coord); cp[mu] = (coord[mu] + 1) % grid[mu]; rp = lu.coord_rank(cp, grid)
            from_fwd, from_bwd = torch.empty_like(send_bwd), torch.empty_like(send_fwd)
            reqs = [dist.isend(send_bwd, rm, tag=2 * mu), dist.isend(send_fwd, rp, tag=2 * mu + 1),
                    dist.irecv(from_fwd, rp, tag=2 * mu), dist.irecv(from_bwd, rm, tag=2 * mu + 1)]
            for r in reqs:
                r.wait()
            f_sh[tuple(hi)] = from_fwd.numpy()
            b_sh[tuple(lo)] = from_bwd.numpy()
        out += np.einsum("...ab,...sb->...sa", Um, f_sh) + b_sh
    return out.reshape(psi.shape)


def _worker(rank, world, port, grid, X, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coord = lu.rank_coord(rank, grid)
        G = tuple(X[d] * grid[d] for d in range(4))
        gam = gamma_ukqcd()
        # global reference (every rank builds it; small)
        Ug = lu.random_su3_lex(G, seed=137)
        xs, ys, zs, ts = lu.coords_lex(G)
        Ug[3, ts == G[3] - 1] *= -1
        psig = lu.r2c(lu.gaussian_spinor_lex(G, seed=101))
        # local shard from the sharded generator (boundary sign only on the last rank in T)
        Ul = lu.random_su3_lex(X, seed=137, grid=grid, coord=coord)
        if coord[3] == grid[3] - 1:
            xl, yl, zl, tl = lu.coords_lex(X)
            Ul[3, tl == X[3] - 1] *= -1
        psil = lu.r2c(lu.gaussian_spinor_lex(X, seed=101, grid=grid, coord=coord))
        gl, _ = lu.global_lex(X, grid, coord)
        errs = []
        for dagger in (False, True):
            ref = lu.dense_hop(Ug, psig, G, gam, dagger=dagger)[gl.astype(np.int64)]
            got = sharded_hop(Ul, psil, X, grid, coord, rank, gam, dagger)
            errs.append(float(np.abs(got - ref).max() / np.abs(ref).max()))
        # scalar all-reduce of the CG: local |psi|^2 summed over ranks equals the global norm
        n = torch.tensor([float(np.sum(np.abs(psil) ** 2))], dtype=torch.float64)
        dist.all_reduce(n)
        errs.append(abs(float(n[0]) - float(np.sum(np.abs(psig) ** 2))) / float(n[0]))
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(1, 1, 1, 2), (1, 1, 2, 1)])
def test_sharded_hop_protocol_world_size_2(grid):
    X = (4, 4, 4, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, grid, X, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-13, (rank, errs)
