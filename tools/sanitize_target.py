"""Target of tools/sanitize.sh (compute-sanitizer): the hot path on 8^3 x 16 -- hop, M^dag M, fp64 and mixed CG, Chebyshev filter,
blas reductions, prepare / reconstruct -- plainly and with the ghost-zone path forced on one GPU in each halo mode
(tmq_force_partition; mode 3 fused compute + peer stores, mode 2 copy-engine peer copies + flag waits, mode 1 peer stores + ticket, mode 0 NCCL-style staging)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")):
    sys.path.insert(0, p)
import numpy as np

import tmq

X = (8, 8, 8, 16)
KAPPA, MU = 1.0 / (2.0 * 4.1), 0.1


def run(part, mode, iters):
    c = tmq.Context(X)
    if part is not None:
        c.force_partition(part)
        c.set_option(tmq.OPT_HALO_P2P, mode)
    gauge = tmq.gen_gauge(X, seed=137, t_boundary=-1)
    c.load_gauge(gauge, t_boundary=-1, recon=12)
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    full = tmq.gen_spinor(X, "gaussian", seed=101)
    Vh = c.Vh
    a, b, x = c.spinor(), c.spinor(), c.spinor()
    a.set(full[:Vh])
    c.dslash(b, a, 1, 0); c.dslash(b, a, 1, 1)
    c.mdagm(b, a)
    n2 = c.norm2(b)
    info = c.cg_mdagm(x, a, tol=1e-30, maxiter=iters)
    infom = c.cg_mdagm(x, a, tol=1e-30, maxiter=iters, sloppy_prec=4, reliable_delta=1e-4)
    c.poly_mdagm(b, a, 3, 0.2, 3.0)
    f, g, s = c.spinor(8, tmq.FULL), c.spinor(8, tmq.FULL), c.spinor()
    f.set(full)
    c.prepare(s, f); c.reconstruct(g, s, f)
    c.mat_full(g, f, 0)
    c.sync()
    print("sanitize target: part=%s mode=%s  |MdagM a|^2 = %.12e  cg iters %d / %d  halo_mode %d" % (part, mode, n2, info["iter"], infom["iter"], c.halo_mode()))
    c.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "plain"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    if what == "plain":
        run(None, 0, iters)
    else:
        mode = int(what[-1])
        for part in ((0, 0, 0, 1), (0, 0, 1, 1)):
            run(part, mode, iters)
