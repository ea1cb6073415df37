#!/usr/bin/env python
"""Tuning sweep on one GPU: Dslash kernel time vs thread->site tile, precision and reconstruct (CUDA events)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT)
import numpy as np
import tmq
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96])
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--tiles", default="4,4,2;2,2,2;8,2,2;4,2,4;2,4,4;1,1,1;8,4,1;4,4,4;8,8,1;2,2,8;16,2,1;4,8,2")
ap.add_argument("--precs", default="8,4")
ap.add_argument("--recons", default="12,18")
a = ap.parse_args()
X = tuple(a.lattice)
Vh = int(np.prod(X)) // 2
gauge = tmq.gen_gauge(X)
src = tmq.gen_spinor(X, "gaussian")[:Vh]
peak, _ = bench.measured_peak()
for recon in [int(r) for r in a.recons.split(",")]:
    c = tmq.Context(X)
    c.load_gauge(gauge, recon=recon)
    c.set_op(bench.KAPPA, bench.MU, 0)
    b = c.spinor(8); b.set(src)
    for prec in [int(p) for p in a.precs.split(",")]:
        for tile in a.tiles.split(";"):
            ty, tz, tt = [int(v) for v in tile.split(",")]
            c.set_tile(ty, tz, tt)
            row = {"recon": recon, "prec": prec, "tile": [ty, tz, tt]}
            for kind in (1, 4):
                ms, _ = c.time_kernel(kind, prec, a.reps, b)
                if kind == 1:
                    gbs = bench.bytes_per_site(1, prec, recon) * Vh / ms * 1e-6
                else:
                    gbs = bench.step_bytes_per_site(prec, recon) * Vh / ms * 1e-6
                row["k%d_ms" % kind] = round(ms, 4); row["k%d_frac" % kind] = round(gbs / peak, 4)
            print(json.dumps(row), flush=True)
    c.close()
