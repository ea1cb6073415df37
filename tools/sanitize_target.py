"""The hot path on 8^3 x 16 -- hop, M^dag M, M_pc, fp64 and mixed CG, Chebyshev filter, prepare / reconstruct, the full operator, blas
reductions, the containers' ghost exchange and plaquette -- plainly and with the ghost-zone path forced on one GPU in each halo mode
(tmq_force_partition; mode 4 fused pack + copy-engine push, mode 3 fused compute + peer stores, mode 2 copy-engine peer copies + flag waits, mode 1 peer stores + ticket,
mode 0 NCCL-style staging), every result checked against the CPU oracle.

Two uses: (1) the target of tools/sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck / initcheck) where the tool is
available; (2) under TMQ_GUARD_BYTES=4096 (tests/test_gpu_guard.py) the library's own out-of-bounds net: red zones around every device
allocation, NaN-filled, so that an out-of-bounds or uninitialised read breaks the parity checks below and tmq.guard_check() reports
every out-of-bounds write."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")):
    sys.path.insert(0, p)
import numpy as np

import lattice_util as lu
import tmq
from oracle.oracle import Oracle

X = (8, 8, 8, 16)
KAPPA, MU = 1.0 / (2.0 * 4.1), 0.1


def run(part, mode, iters, o, gauge, full):
    c = tmq.Context(X)
    if part is not None:
        c.force_partition(part)
        c.set_option(tmq.OPT_HALO_P2P, mode)
    c.load_gauge(gauge, t_boundary=-1, recon=12)
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    Vh = c.Vh
    even, odd = np.ascontiguousarray(full[:Vh]), np.ascontiguousarray(full[Vh:])
    worst = 0.0
    for prec, tol in ((8, 1e-13), (4, 1e-5)):
        a, b = c.spinor(prec), c.spinor(prec)
        a.set(even)
        for dag in (0, 1):
            c.dslash(b, a, 1, dag)
            worst = max(worst, lu.rel_l2(b.get(), o.dslash(gauge, even, 1, dag)) / tol)
            c.matpc(b, a, dag)
            worst = max(worst, lu.rel_l2(b.get(), o.matpc(gauge, even, KAPPA, MU, 0, dag)) / (2 * tol))
        c.mdagm(b, a)
        worst = max(worst, lu.rel_l2(b.get(), o.mdagm(gauge, even, KAPPA, MU, 0)) / (2 * tol))
        c.poly_mdagm(b, a, 3, 0.2, 3.0)
        assert np.isfinite(b.get()).all()
        a.free(); b.free()
    a, x = c.spinor(), c.spinor()
    a.set(even)
    x_ref, it_ref, _, hist = o.cg_mdagm(gauge, even, KAPPA, MU, 0, tol=1e-30, maxiter=iters)
    info = c.cg_mdagm(x, a, tol=1e-30, maxiter=iters)
    assert np.allclose(c.cg_history(iters + 1), hist[: iters + 1], rtol=1e-9), "CG history"
    worst = max(worst, lu.rel_l2(x.get(), x_ref) / 1e-10)
    infom = c.cg_mdagm(x, a, tol=1e-8, maxiter=500, sloppy_prec=4, reliable_delta=1e-4)
    assert infom["true_res"] <= 1.05e-8
    f, g, s = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL), c.spinor()
    f.set(full)
    c.prepare(s, f); c.reconstruct(g, s, f)
    assert np.isfinite(g.get()).all()
    c.mat_full(g, f, 0)
    worst = max(worst, lu.rel_l2(g.get(), o.mat(gauge, full, KAPPA, MU, 0)) / 1e-13)
    n2 = c.norm2(f)
    assert abs(n2 - float(np.sum(full * full))) < 1e-12 * n2
    # the containers' ghost zones and the plaquette read through them
    V = 2 * Vh
    U = lu.r2c(np.stack([lu.spinor_lex_from_eo(gauge[mu], X) for mu in range(4)]))
    gq = lu.c2r(np.ascontiguousarray(np.transpose(U, (0, 2, 3, 1)))).reshape(36, V, 2)
    ng = c.qkxtm_ghost_sites()
    buf = np.concatenate([gq.reshape(-1), np.zeros(ng * 72)])
    p = c.dev_malloc(buf.nbytes)
    c.h2d(p, buf)
    plaq = c.qkxtm_plaquette(p, 8)
    assert abs(plaq - o.plaquette(gauge)) < 1e-12, (plaq, o.plaquette(gauge))
    c.dev_free(p)
    c.sync()
    assert worst < 1.0, worst
    print("sanitize target: part=%s mode=%s  worst error / tolerance %.3f  cg iters %d / mixed %d  halo_mode %d" % (part, mode, worst, info["iter"], infom["iter"], c.halo_mode()))
    c.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "plain"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    o = Oracle(X)
    gauge = tmq.gen_gauge(X, seed=137, t_boundary=-1)
    full = tmq.gen_spinor(X, "gaussian", seed=101)
    for w in (["plain", "mode4", "mode3", "mode2", "mode1", "mode0"] if what == "all" else [what]):
        if w == "plain":
            run(None, 0, iters, o, gauge, full)
        else:
            for part in ((0, 0, 0, 1), (0, 0, 1, 0), (0, 0, 1, 1)):
                run(part, int(w[-1]), iters, o, gauge, full)
    bad, msg = tmq.guard_check()
    print("guard check: %d corrupted allocations %s (TMQ_GUARD_BYTES=%s)" % (bad, msg, os.environ.get("TMQ_GUARD_BYTES", "unset")))
    sys.exit(1 if bad else 0)
