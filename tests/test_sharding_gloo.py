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


# ---- the container layer on a process grid: smearing with exchanged z faces, contraction with the zero-padded global-T all-reduce ----
def _slab(a, G, X, coord):
    """local block of a global lexicographic array [...][Vg] -> [...][Vl]"""
    g = a.reshape(a.shape[:-1] + (G[3], G[2], G[1], G[0]))
    sl = tuple(slice(coord[d] * X[d], (coord[d] + 1) * X[d]) for d in (3, 2, 1, 0))
    return np.ascontiguousarray(g[(Ellipsis,) + sl]).reshape(a.shape[:-1] + (-1,))


def _sendrecv(send_bwd, send_fwd, rm, rp, tag):
    """send_bwd -> rank-1, send_fwd -> rank+1; returns (from rank+1, from rank-1): csrc/tmq_comm.cpp comm_sendrecv_dim"""
    a, b = torch.from_numpy(np.ascontiguousarray(send_bwd)), torch.from_numpy(np.ascontiguousarray(send_fwd))
    from_fwd, from_bwd = torch.empty_like(a), torch.empty_like(b)
    reqs = [dist.isend(a, rm, tag=tag), dist.isend(b, rp, tag=tag + 1), dist.irecv(from_fwd, rp, tag=tag), dist.irecv(from_bwd, rm, tag=tag + 1)]
    for r in reqs:
        r.wait()
    return from_fwd.numpy(), from_bwd.numpy()


def _smear_step_sharded(vec, gauge, X, grid, coord, alpha):
    """one Gaussian smearing step on a slab (csrc/tmq_smear.cu): x, y hop locally; z takes the neighbours' faces when z is split (the psi
    faces every step, the backward neighbour's U_z face as well -- the product exchanges that one once per call)"""
    Xd, Yd, Zd, Td = X
    psi = vec.reshape(4, 3, Td, Zd, Yd, Xd); U = gauge.reshape(4, 3, 3, Td, Zd, Yd, Xd)
    acc = np.zeros_like(psi)
    for mu, ax in ((0, 5), (1, 4), (2, 3)):
        fwd = np.roll(psi, -1, axis=ax); bwd = np.roll(psi, 1, axis=ax); Ub = np.roll(U[mu], 1, axis=ax)
        if mu == 2 and grid[2] > 1:
            cm = list(coord); cm[2] = (coord[2] - 1) % grid[2]; cp = list(coord); cp[2] = (coord[2] + 1) % grid[2]
            rm, rp = lu.coord_rank(cm, grid), lu.coord_rank(cp, grid)
            lo, hi = psi[:, :, :, :1], psi[:, :, :, Zd - 1:]
            from_fwd, from_bwd = _sendrecv(lo, hi, rm, rp, 10)
            fwd[:, :, :, Zd - 1:] = from_fwd; bwd[:, :, :, :1] = from_bwd
            _, u_from_bwd = _sendrecv(U[2][:, :, :, Zd - 1:], U[2][:, :, :, Zd - 1:], rm, rp, 20)
            Ub[:, :, :, :1] = u_from_bwd
        acc += np.einsum("ab...,sb...->sa...", U[mu], fwd) + np.einsum("ba...,sb...->sa...", Ub.conj(), bwd)
    return ((psi + alpha * acc) / (1.0 + 6.0 * alpha)).reshape(12, -1)


def _worker_containers(rank, world, port, grid, X, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        coord = lu.rank_coord(rank, grid)
        G = tuple(X[d] * grid[d] for d in range(4))
        Vg = int(np.prod(G))
        rng = np.random.default_rng(7)                        # the same global fields on every rank
        vec_g = rng.standard_normal((12, Vg)) + 1j * rng.standard_normal((12, Vg))
        U_g = rng.standard_normal((4, 3, 3, Vg)) + 1j * rng.standard_normal((4, 3, 3, Vg))
        errs = []
        # smearing: 3 steps on the slab with exchanged faces = slab of 3 global steps
        cur = _slab(vec_g, G, X, coord); Ul = _slab(U_g, G, X, coord)
        for _ in range(3):
            cur = _smear_step_sharded(cur, Ul, X, grid, coord, 4.0)
        want = _slab(O.gauss_smear(vec_g, U_g, G, 4.0, 3), G, X, coord)
        errs.append(float(np.abs(cur - want).max() / np.abs(want).max()))
        # meson contraction: local site values, phases from GLOBAL coordinates, own time slices of a zero-padded global-T buffer, one all-reduce
        p1_g = (rng.standard_normal((144, Vg)) + 1j * rng.standard_normal((144, Vg))).reshape(4, 4, 3, 3, Vg)
        p2_g = (rng.standard_normal((144, Vg)) + 1j * rng.standard_normal((144, Vg))).reshape(4, 4, 3, 3, Vg)
        moms, src = O.create_momenta(2), (1, 2, G[2] - 1)
        c = np.stack([O.contract_mesons_site(_slab(p1_g, G, X, coord)), O.contract_mesons_site(_slab(p2_g, G, X, coord))])
        c = c.reshape(2, 10, X[3], X[2], X[1], X[0])
        x = np.arange(X[0]) + coord[0] * X[0] - src[0]; y = np.arange(X[1]) + coord[1] * X[1] - src[1]; z = np.arange(X[2]) + coord[2] * X[2] - src[2]
        buf = np.zeros((G[3], len(moms), 2, 10), dtype=np.complex128)
        for im, (px, py, pz) in enumerate(moms):
            ph = np.exp(-2j * np.pi * (pz * z[:, None, None] / G[2] + py * y[None, :, None] / G[1] + px * x[None, None, :] / G[0]))
            buf[coord[3] * X[3]:(coord[3] + 1) * X[3], im] = np.einsum("uptzyx,zyx->tup", c, ph)
        t = torch.from_numpy(np.ascontiguousarray(np.stack([buf.real, buf.imag], axis=-1)))
        dist.all_reduce(t)                                    # sums the z ranks and gathers the t ranks (csrc/tmq_contract.cu: allreduce_host)
        got = t.numpy()[..., 0] + 1j * t.numpy()[..., 1]
        want = O.contract_mesons_mom(p1_g, p2_g, G, moms, src)
        errs.append(float(np.abs(got - want).max() / np.abs(want).max()))
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(1, 1, 1, 2), (1, 1, 2, 1)])
def test_container_layer_protocols_world_size_2(grid):
    """Gaussian smearing with exchanged z faces and the contraction's zero-padded global-T all-reduce, over gloo on 2 CPU ranks"""
    X = (4, 4, 4, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_containers, args=(r, 2, port, grid, X, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-12, (rank, errs)
