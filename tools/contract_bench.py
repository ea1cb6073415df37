#!/usr/bin/env python
"""Timing of the meson two-point contraction (csrc/tmq_contract.cu) on one B200: the site kernel reads 2 x 144 complex per site
and writes 20 complex doubles (algorithmic bytes = 288 * 2 * sizeof(real) + 320 per site), then the separable Fourier sum.
The reference runs one launch + one blocking D2H per time slice and projects every momentum with a 64-thread shared-memory
tree (lib/qudaQKXTM_kernels.cu:1127-1199)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import tmq  # noqa: E402
from oracle.oracle import create_momenta  # noqa: E402  (momentum list only)

ap = argparse.ArgumentParser()
ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96])
ap.add_argument("--precs", default="4,8")
ap.add_argument("--qsq", default="0,3,16")
ap.add_argument("--baryons", type=int, default=0, help="also time the baryon contraction (Q_sq = 3)")
a = ap.parse_args()
X = tuple(a.lattice)
V = int(np.prod(X))
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
c = tmq.Context(X)
for prec in [int(p) for p in a.precs.split(",")]:
    nb = V * 288 * prec
    p1, p2 = c.dev_malloc(nb), c.dev_malloc(nb)
    for p in (p1, p2):
        c.L.tmq_dev_memset(c.h, p, 0x3c, nb)            # small normal numbers (timing only; parity is tests/test_gpu_contract.py)
    for q in [int(v) for v in a.qsq.split(",")]:
        moms = create_momenta(q)
        c.qkxtm_contract_mesons(p1, p2, prec, moms, (1, 2, 3), global_T=X[3])     # warm-up
        l0 = c.launch_count()
        t0 = time.perf_counter()
        c.timer_start()
        c.qkxtm_contract_mesons(p1, p2, prec, moms, (1, 2, 3), global_T=X[3])
        ms = c.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        gb = V * (288 * 2 * prec + 320) * 1e-9
        print(json.dumps({"what": "meson contraction (10 channels x 2 propagators) + momentum projection", "lattice": X, "prec": prec,
                          "Q_sq": q, "nmoms": len(moms), "device_span_ms": ms, "wall_ms": wall, "launches": c.launch_count() - l0,
                          "site_kernel_algorithmic_GB": gb, "GB/s_if_all_time_were_the_site_kernel": gb / (ms * 1e-3),
                          "frac_of_hbm_peak": gb / (ms * 1e-3) / peak}), flush=True)
    if a.baryons:
        moms = create_momenta(3)
        c.qkxtm_contract_baryons(p1, p2, prec, moms, (1, 2, 3), X[3])           # warm-up (work space)
        l0 = c.launch_count()
        c.timer_start()
        c.qkxtm_contract_baryons(p1, p2, prec, moms, (1, 2, 3), X[3])
        ms = c.timer_stop()
        print(json.dumps({"what": "baryon contraction (10 channels x 4x4 spin x 2 assignments) + momentum projection", "lattice": X, "prec": prec,
                          "Q_sq": 3, "nmoms": len(moms), "device_span_ms": ms, "launches": c.launch_count() - l0}), flush=True)
    for p in (p1, p2):
        c.dev_free(p)
c.close()
