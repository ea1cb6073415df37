"""debug: fused halo mode on one GPU with a forced partition, step by step (short halo timeout)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")):
    sys.path.insert(0, p)
import numpy as np
import tmq
X = (8, 8, 8, 16)
KAPPA, MU = 1.0 / (2.0 * 4.1), 0.1
gauge = tmq.gen_gauge(X, seed=137, t_boundary=-1)
full = tmq.gen_spinor(X, "gaussian", seed=101)
for part in ((0, 0, 0, 1), (0, 0, 1, 0), (0, 0, 1, 1)):
    c = tmq.Context(X)
    c.force_partition(part)
    c.set_option(tmq.OPT_HALO_P2P, 3)
    c.set_option(7, 3000)
    c.load_gauge(gauge, t_boundary=-1, recon=12)
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    a, b, x = c.spinor(), c.spinor(), c.spinor()
    a.set(full[:c.Vh])
    def step(name, fn):
        try:
            r = fn(); c.sync(); print(part, name, "ok", r if isinstance(r, dict) else "", flush=True)
        except Exception as e:
            print(part, name, "FAILED", e, flush=True)
    step("dslash", lambda: c.dslash(b, a, 1, 0))
    step("mdagm", lambda: c.mdagm(b, a))
    step("matpc", lambda: c.matpc(b, a, 0))
    for it in (1, 2, 3, 10):
        step("cg maxiter %d" % it, lambda: c.cg_mdagm(x, a, tol=1e-30, maxiter=it))
    step("cg mixed", lambda: c.cg_mdagm(x, a, tol=1e-8, maxiter=200, sloppy_prec=4))
    step("mdagm after", lambda: c.mdagm(b, a))
    c.close()
