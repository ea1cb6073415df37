#!/usr/bin/env python
"""A/B of tuning options on one GPU (CUDA events): per-kernel and fused-iteration times."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT)
import numpy as np, tmq, bench
X = (48, 48, 48, 96); Vh = int(np.prod(X)) // 2
c = tmq.Context(X); c.load_gauge(tmq.gen_gauge(X), recon=12); c.set_op(bench.KAPPA, bench.MU, 0)
b = c.spinor(8); b.set(tmq.gen_spinor(X, "gaussian")[:Vh])
for rep in range(2):
    for pf in (0, 1):
        c.set_option(1, pf)
        row = {"prefetch": pf}
        for prec in (8, 4):
            for kind in (1, 2, 3, 4):
                ms, _ = c.time_kernel(kind, prec, 20, b)
                row["p%d_k%d" % (prec, kind)] = round(ms, 4)
        print(json.dumps(row), flush=True)
