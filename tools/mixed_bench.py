#!/usr/bin/env python
"""Mixed-precision CG (fp32 inner iterations, fp64 reliable updates) vs pure fp64 on one B200: solve time, iterations and true
residual as a function of reliable_delta (the reference drivers set 1e-4, qkxtm/Calc_Loops.cpp:481), on the bench operator
and on a worse-conditioned one."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import tmq  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96])
ap.add_argument("--tol", type=float, default=1e-9)
a = ap.parse_args()
X = tuple(a.lattice)
Vh = int(np.prod(X)) // 2
c = tmq.Context(X)
c.load_gauge(tmq.gen_gauge(X), t_boundary=-1, recon=12)
b = c.spinor(8); b.set(tmq.gen_spinor(X, "z4")[:Vh])
x = c.spinor(8)
for kappa, mu in ((1.0 / (2.0 * 4.1), 0.1), (0.14, 0.005), (0.15, 0.002)):
    c.set_op(kappa, mu, 0)
    c.cg_mdagm(x, b, tol=1e-3, maxiter=50)            # warm-up
    r = c.cg_mdagm(x, b, tol=a.tol, maxiter=20000)
    print(json.dumps({"kappa": kappa, "mu": mu, "prec": "fp64", "iter": r["iter"], "true_res": r["true_res"], "secs": r["secs"]}), flush=True)
    for delta in (1e-1, 1e-2, 1e-3, 1e-4):
        m = c.cg_mdagm(x, b, tol=a.tol, maxiter=20000, sloppy_prec=4, reliable_delta=delta)
        print(json.dumps({"kappa": kappa, "mu": mu, "prec": "fp32/fp64", "reliable_delta": delta, "iter": m["iter"], "true_res": m["true_res"],
                          "secs": m["secs"], "speedup_vs_fp64": r["secs"] / m["secs"]}), flush=True)
c.close()
