#!/usr/bin/env python
"""Single-GPU cost of the ghost-zone machinery: unpartitioned vs forced self-partition (peer-memory / copy path)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT)
import numpy as np, tmq, bench
for X in [(48, 48, 48, 12), (48, 48, 48, 96)]:
    Vh = int(np.prod(X)) // 2
    gauge = tmq.gen_gauge(X); src = tmq.gen_spinor(X, "gaussian")[:Vh]
    for mode in ("none", "p2p", "p2p0", "p2p100", "copy"):
        c = tmq.Context(X)
        if mode != "none":
            c.force_partition((0, 0, 0, 1))
            c.set_option(tmq.OPT_HALO_P2P, 0 if mode == "copy" else 1)
            if mode == "p2p0": c.set_option(3, 0)
            if mode == "p2p100": c.set_option(3, 100)
        c.load_gauge(gauge, recon=12); c.set_op(bench.KAPPA, bench.MU, 0)
        b = c.spinor(8); b.set(src)
        row = {"X": X, "mode": mode}
        for rep in range(2):
            for kind in (1, 4):
                ms, nl = c.time_kernel(kind, 8, 30, b)
                row["k%d_ms_%d" % (kind, rep)] = round(ms, 4); row["k%d_launches" % kind] = nl
        print(json.dumps(row), flush=True)
        c.close()
