#!/usr/bin/env python
"""Timing of the twisted-clover variant on one B200: hop + A^-1 kernel and the fused CG iteration with
the site-dependent 6x6 chiral blocks.  Algorithmic bytes per parity site: twisted-mass bytes + 72 complex = 1152 B (fp64) for
every launch that applies A^-1 (K1, K2, K3); K2 also stores y = M p, which K4 takes as its x term."""
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
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
X = tuple(a.lattice)
Vh = int(np.prod(X)) // 2
KAPPA, MU, CSW = 1.0 / (2.0 * 4.1), 0.1, 1.57551
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
c = tmq.Context(X)
c.load_gauge(tmq.gen_gauge(X), t_boundary=-1, recon=12)
c.set_op(KAPPA, MU, 0)
b = c.spinor(8); b.set(tmq.gen_spinor(X, "z4")[:Vh])
for clover in (0, 1):
    if clover:
        c.timer_start()
        c.clover_load(CSW * KAPPA)
        print(json.dumps({"what": "clover field + inverse construction", "lattice": X, "ms": c.timer_stop()}), flush=True)
    for prec in (8, 4):
        ms1, _ = c.time_kernel(1, prec, a.reps, b)
        ms4, n4 = c.time_kernel(4, prec, a.reps, b)
        cl = 72 * 2 * prec * clover          # one 2 x 6x6 complex field per site
        b1 = (24 + 24 + 8 * 12) * prec + cl
        # K1, K2, K3 load A^-1; K2 also stores y = M p and K4 reads it instead of w (one extra 24-real write)
        b4 = (24 * 16 + 32 * 12) * prec + clover * (3 * cl + 24 * prec)
        print(json.dumps({"what": "twisted-clover" if clover else "twisted-mass", "lattice": X, "prec": prec,
                          "hop+Ainv_ms": ms1, "hop+Ainv_bytes_per_site": b1, "hop+Ainv_frac_of_hbm_peak": b1 * Vh / ms1 * 1e-6 / peak,
                          "cg_iter_ms": ms4, "cg_iter_launches": n4, "cg_iter_bytes_per_site": b4,
                          "cg_iter_frac_of_hbm_peak": b4 * Vh / ms4 * 1e-6 / peak}), flush=True)
c.close()
