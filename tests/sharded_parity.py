"""Sharded (multi-GPU) parity check, one process per GPU.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2
Every rank computes the CPU oracle on the GLOBAL lattice and compares its own slab of the device result
(hop both parities / daggers, M^dag M, CG iterations, true residual).  Exit code 0 = parity green."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import lattice_util as lu
import tmq
from oracle.oracle import Oracle

KAPPA, MU = 1.0 / (2.0 * 4.1), 0.1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=4, default=[8, 8, 8, 16])
    ap.add_argument("--grid", type=int, nargs=4, default=[1, 1, 1, 2])
    ap.add_argument("--recon", type=int, default=12)
    ap.add_argument("--p2p", type=int, default=4)
    ap.add_argument("--pack-async", type=int, default=0)
    ap.add_argument("--clover", type=int, default=1, help="also check the twisted-clover variant")
    ap.add_argument("--amin", type=float, default=0.2, help="lower edge of the Chebyshev window (just above the wanted eigenvalues)")
    ap.add_argument("--eig", type=int, default=1, help="also run the (slower) eigensolver / deflation check")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, grid = tuple(a.lattice), tuple(a.grid)
    assert int(np.prod(grid)) == world
    X = tuple(G[d] // grid[d] for d in range(4))
    coord = lu.rank_coord(rank, grid)
    ctx = tmq.Context(X, grid=grid, coord=coord, device=int(os.environ.get("LOCAL_RANK", rank)))
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(tmq.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, src=0)
    ctx.comm_init(uid.numpy().tobytes(), world, rank)
    ctx.set_option(tmq.OPT_HALO_P2P, a.p2p)
    if a.pack_async:
        ctx.set_option(5, 1)
    mode = ctx.halo_mode()
    ctx.load_gauge(tmq.gen_gauge(X, grid=grid, coord=coord), t_boundary=-1, recon=a.recon)
    ctx.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)

    o = Oracle(G)
    gauge_g = tmq.gen_gauge(G)
    psi_g = tmq.gen_spinor(G)                                   # global, even-odd order
    Vh_g, Vh = o.Vh, ctx.Vh
    loc = lu.local_from_global_eo(psi_g, X, grid, coord)        # local FULL field
    assert np.array_equal(loc, tmq.gen_spinor(X, grid=grid, coord=coord)), "sharded generator != slab of global field"
    fails = []
    for prec, tol in ((8, 1e-13), (4, 1e-5)):
        s_in, s_out = ctx.spinor(prec), ctx.spinor(prec)
        for out_parity in (0, 1):
            src_g = np.ascontiguousarray(psi_g[(1 - out_parity) * Vh_g:(2 - out_parity) * Vh_g])
            s_in.set(loc[(1 - out_parity) * Vh:(2 - out_parity) * Vh])
            for dagger in (0, 1):
                ref = np.zeros_like(psi_g)
                ref[out_parity * Vh_g:(out_parity + 1) * Vh_g] = o.dslash(gauge_g, src_g, out_parity, dagger)
                ref_loc = lu.local_from_global_eo(ref, X, grid, coord)[out_parity * Vh:(out_parity + 1) * Vh]
                ctx.dslash(s_out, s_in, out_parity, dagger)
                e = lu.rel_l2(s_out.get(), ref_loc)
                if not e < tol:
                    fails.append(("hop", prec, out_parity, dagger, e))
        even_g = np.ascontiguousarray(psi_g[:Vh_g])
        ref = np.zeros_like(psi_g); ref[:Vh_g] = o.mdagm(gauge_g, even_g, KAPPA, MU, 0)
        s_in.set(loc[:Vh]); ctx.mdagm(s_out, s_in)
        e = lu.rel_l2(s_out.get(), lu.local_from_global_eo(ref, X, grid, coord)[:Vh])
        if not e < 2 * tol:
            fails.append(("mdagm", prec, e))
    # CG: iteration count and residual against the CPU CG on the global lattice
    x_ref, it_ref, tr_ref, _ = o.cg_mdagm(gauge_g, even_g, KAPPA, MU, 0, tol=1e-9, maxiter=5000)
    b, x = ctx.spinor(8), ctx.spinor(8)
    b.set(loc[:Vh])
    info = ctx.cg_mdagm(x, b, tol=1e-9, maxiter=5000)
    xr = np.zeros_like(psi_g); xr[:Vh_g] = x_ref
    e = lu.rel_l2(x.get(), lu.local_from_global_eo(xr, X, grid, coord)[:Vh])
    if abs(info["iter"] - it_ref) > 2 or info["true_res"] > 1.05e-9 or e > 1e-8:
        fails.append(("cg", info, it_ref, e))
    info4 = ctx.cg_mdagm(x, b, tol=1e-9, maxiter=5000, sloppy_prec=4, reliable_delta=0.1)
    e4 = lu.rel_l2(x.get(), lu.local_from_global_eo(xr, X, grid, coord)[:Vh])
    if info4["true_res"] > 1.05e-9 or e4 > 1e-7:
        fails.append(("cg-mixed", info4, e4))
    # eigensolver layer on the sharded lattice: Chebyshev filter vs the oracle recurrence, a few eigenpairs vs ARPACK on
    # the global oracle operator, and the deflation projector built from them (coefficients all-reduced over ranks)
    from oracle.oracle import poly_operator, eigs_reference
    cplx = lambda v: np.ascontiguousarray(v[..., 0] + 1j * v[..., 1]).ravel()
    real = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(Vh_g, 4, 3, 2))
    A = lambda v: cplx(o.mdagm(gauge_g, real(v), KAPPA, MU, 0))
    slab = lambda v: lu.local_from_global_eo(np.concatenate([real(v), np.zeros((Vh_g, 4, 3, 2))]), X, grid, coord)[:Vh]
    pin, pout = ctx.spinor(8), ctx.spinor(8)
    pin.set(loc[:Vh])
    ctx.poly_mdagm(pout, pin, 12, 0.3, 2.0)
    e = lu.rel_l2(pout.get(), slab(poly_operator(A, cplx(even_g), 12, 0.3, 2.0)))
    if not e < 1e-12:
        fails.append(("cheb", e))
    if a.eig:
        nev, nkv = 4, 24
        lam_ref, U_ref = eigs_reference(A, 12 * Vh_g, nev, nkv, "SR", poly=(20, a.amin, 2.0), tol=1e-12)
        es = ctx.eigset(nkv + 1)
        r = ctx.eigensolve(es, nev, nkv, poly_deg=20, amin=a.amin, amax=2.0, tol=1e-11, max_restarts=500, which=0, seed=9)
        if r["nconv"] != nev or not np.allclose(r["evals"], lam_ref, rtol=1e-9) or r["resid"].max() > 1e-8:
            fails.append(("eig", r, lam_ref))
        ctx.deflate(pout, pin, es, r["evals"], nev)
        want = U_ref @ ((U_ref.conj().T @ cplx(even_g)) / lam_ref)
        e = lu.rel_l2(pout.get(), slab(want))
        if not e < 1e-7:
            fails.append(("deflate", e))
    # twisted-clover on the sharded lattice: the clover term is built from the gauge field with an exchanged halo
    if a.clover:
        csw = 1.57551
        ctx.clover_load(csw * KAPPA)
        o.set_clover(o.clover_compute(gauge_g, csw * KAPPA))
        ref = np.zeros_like(psi_g); ref[:Vh_g] = o.mdagm(gauge_g, even_g, KAPPA, MU, 0)
        s_in, s_out = ctx.spinor(8), ctx.spinor(8)
        s_in.set(loc[:Vh]); ctx.mdagm(s_out, s_in)
        e = lu.rel_l2(s_out.get(), lu.local_from_global_eo(ref, X, grid, coord)[:Vh])
        if not e < 4e-13:
            fails.append(("clover-mdagm", e))
        x_ref, it_ref_c, _, _ = o.cg_mdagm(gauge_g, even_g, KAPPA, MU, 0, tol=1e-9, maxiter=5000)
        info_c = ctx.cg_mdagm(x, b, tol=1e-9, maxiter=5000)
        if abs(info_c["iter"] - it_ref_c) > 2 or info_c["true_res"] > 1.05e-9:
            fails.append(("clover-cg", info_c, it_ref_c))
        o.set_clover(None)
        ctx.clover_free()
    # container layer on the sharded lattice (QKXTM device layouts, x lexicographic): Gaussian smearing (z faces exchanged before
    # every step; nothing to exchange on a t split) and the meson contraction (z ranks summed and t ranks gathered by one
    # all-reduce) against the oracle's restatements on the GLOBAL lattice
    from oracle.oracle import gauss_smear, contract_mesons_mom, create_momenta
    rng = np.random.default_rng(41)
    Vg = int(np.prod(G)); Vl = int(np.prod(X))
    gshape = (G[3], G[2], G[1], G[0])
    sl = tuple(slice(coord[d] * X[d], (coord[d] + 1) * X[d]) for d in (3, 2, 1, 0))
    loc_lex = lambda f: np.ascontiguousarray(f.reshape(f.shape[:-1] + gshape)[(Ellipsis,) + sl]).reshape(f.shape[:-1] + (Vl,))
    c2r = lambda f: np.ascontiguousarray(np.stack([f.real, f.imag], axis=-1))
    r2c = lambda f: f[..., 0] + 1j * f[..., 1]
    def put(arr):
        ptr = ctx.dev_malloc(arr.nbytes); ctx.h2d(ptr, np.ascontiguousarray(arr)); return ptr
    vec_g = rng.standard_normal((12, Vg)) + 1j * rng.standard_normal((12, Vg))
    Ulex = lu.random_su3_lex(G, seed=5)                                      # [4][x_lex][3][3]
    gq_g = np.ascontiguousarray(np.transpose(Ulex, (0, 2, 3, 1))).reshape(36, Vg)
    nsm, alpha = 5, 4.0
    want = loc_lex(gauss_smear(vec_g, gq_g.reshape(4, 3, 3, Vg), G, alpha, nsm))
    d_in, d_g = put(c2r(loc_lex(vec_g))), put(c2r(loc_lex(gq_g)))
    d_out = put(np.zeros((12, Vl, 2)))
    ctx.qkxtm_gauss_smear(d_out, d_in, d_g, 8, nsm, alpha)
    got = np.empty((12, Vl, 2)); ctx.d2h(got, d_out)
    e = lu.rel_l2(r2c(got), want)
    if not e < 1e-13:
        fails.append(("gauss-smear", e))
    p1_g = rng.standard_normal((144, Vg)) + 1j * rng.standard_normal((144, Vg))
    p2_g = rng.standard_normal((144, Vg)) + 1j * rng.standard_normal((144, Vg))
    moms, srcp = create_momenta(2), (1, G[1] - 1, G[2] - 2)
    want = contract_mesons_mom(p1_g.reshape(4, 4, 3, 3, Vg), p2_g.reshape(4, 4, 3, 3, Vg), G, moms, srcp)
    d_p1, d_p2 = put(c2r(loc_lex(p1_g))), put(c2r(loc_lex(p2_g)))
    mom, _ = ctx.qkxtm_contract_mesons(d_p1, d_p2, 8, moms, srcp, global_T=G[3])
    e = np.abs(mom - want).max() / np.abs(want).max()
    if not e < 1e-12:
        fails.append(("contract-mesons", e))
    from oracle import ref as qref
    if qref.available():                                                     # the reference's own baryon kernel body on the global lattice
        bm = [(0, 0, 0), (1, 0, -1)]
        wantb = qref.Ref(G).contract_baryons_mom(c2r(p1_g.reshape(4, 4, 3, 3, Vg)), c2r(p2_g.reshape(4, 4, 3, 3, Vg)), bm, srcp)
        wantb = wantb[..., 0] + 1j * wantb[..., 1]
        gotb = ctx.qkxtm_contract_baryons(d_p1, d_p2, 8, bm, srcp, G[3])
        e = np.abs(gotb - wantb).max() / np.abs(wantb).max()
        if not e < 1e-12:
            fails.append(("contract-baryons", e))
    for ptr in (d_in, d_g, d_out, d_p1, d_p2):
        ctx.dev_free(ptr)
    n2 = ctx.norm2(b)
    if abs(n2 - np.sum(even_g * even_g)) > 1e-12 * n2:
        fails.append(("norm2-allreduce", n2))
    print("halo_mode %d rank %d/%d coord %s: cg iters %d (cpu %d) true_res %.2e mixed iters %d; failures: %s"
          % (mode, rank, world, coord, info["iter"], it_ref, info["true_res"], info4["iter"], fails), flush=True)
    t = torch.tensor([len(fails)], dtype=torch.int64)
    dist.all_reduce(t)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
