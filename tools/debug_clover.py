#!/usr/bin/env python
"""which halo mode / CG look-ahead breaks the clover CG on a forced Z partition (debug aid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, tmq
import lattice_util as lu
from oracle.oracle import Oracle
X = (4, 6, 4, 8); KAPPA = 0.12195121951219513; MU = 0.1; CSW = 1.57551
orc = Oracle(X)
gauge = lu.random_gauge_qdp(X, seed=137, t_boundary=-1)
clov = orc.clover_compute(gauge, CSW * KAPPA)
full = lu.spinor_eo_from_lex(lu.gaussian_spinor_lex(X, seed=101), X)
even = np.ascontiguousarray(full[: orc.Vh])
orc.set_clover(clov)
_, it_ref, _, _ = orc.cg_mdagm(gauge, even, KAPPA, MU, 0, tol=1e-9, maxiter=2000)
for part in ((0, 0, 1, 0), (0, 0, 1, 1), (0, 0, 0, 1)):
    for clover in (1, 0):
        for p2p in (4, 3, 2, 0):
            for lag in (1, 0):
                c = tmq.Context(X)
                c.force_partition(part)
                c.set_option(tmq.OPT_HALO_P2P, p2p); c.set_option(tmq.OPT_CG_LAG, lag)
                c.load_gauge(gauge, t_boundary=-1, recon=12)
                c.set_op(KAPPA, MU, 0)
                if clover: c.clover_load(CSW * KAPPA)
                a, x = c.spinor(), c.spinor()
                a.set(even)
                try:
                    info = c.cg_mdagm(x, a, tol=1e-9, maxiter=300)
                    print("part", part, "clover", clover, "p2p", p2p, "lag", lag, "iters", info["iter"], "(cpu clover %d)" % it_ref, "true_res %.2e" % info["true_res"], flush=True)
                except Exception as e:
                    print("part", part, "clover", clover, "p2p", p2p, "lag", lag, "ERROR", str(e)[:100], flush=True)
                c.close()
