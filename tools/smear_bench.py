#!/usr/bin/env python
"""Timing of the Gaussian smearing on one B200: ms per smearing step for the plain streaming order
and for the L2-resident blocks of time slices; algorithmic bytes = (24 + 24 + 54) * sizeof(real) per site and step."""
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
ap.add_argument("--nsmear", type=int, default=50)
ap.add_argument("--blocks", default="1000000,0,1,2,4")
ap.add_argument("--precs", default="8,4")
a = ap.parse_args()
X = tuple(a.lattice)
V = int(np.prod(X))
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
c = tmq.Context(X)
for prec in [int(p) for p in a.precs.split(",")]:
    nb_vec, nb_g = V * 24 * prec, V * 72 * prec
    din, dout, dg = c.dev_malloc(nb_vec), c.dev_malloc(nb_vec), c.dev_malloc(nb_g)
    for blk in [int(b) for b in a.blocks.split(",")]:
        # small normal numbers everywhere (timing only; parity is tests/test_gpu_smear.py)
        for p, n in ((din, nb_vec), (dout, nb_vec), (dg, nb_g)):
            c.L.tmq_dev_memset(c.h, p, 0x3c, n)
        c.set_option(tmq.OPT_SMEAR_BLOCK_T, blk)
        c.qkxtm_gauss_smear(dout, din, dg, prec, 2, 0.25)       # warm-up
        l0 = c.launch_count()
        c.timer_start()
        c.qkxtm_gauss_smear(dout, din, dg, prec, a.nsmear, 0.25)
        ms = c.timer_stop()
        step = ms / a.nsmear
        gbs = (24 + 24 + 54) * prec * V / (step * 1e-3) * 1e-9
        print(json.dumps({"what": "gaussian smearing", "lattice": X, "prec": prec, "nsmear": a.nsmear,
                          "block_t": "auto" if blk == 0 else ("all" if blk >= X[3] else blk), "ms_total": ms, "ms_per_step": step,
                          "launches": c.launch_count() - l0, "algorithmic_GB/s": gbs, "frac_of_hbm_peak": gbs / peak}), flush=True)
    for p in (din, dout, dg):
        c.dev_free(p)
c.close()
