#!/usr/bin/env python
"""Timing of the eigensolver layer on one B200: the Chebyshev filter per degree at 48^3x96
(roofline: 5376 B per parity site per degree, fp64 recon-12) and one thick-restart Lanczos solve on a smaller
lattice.  Prints one JSON line per measurement."""
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

KAPPA, MU = 1.0 / (2.0 * 4.1), 0.1


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def cheb(X, prec, recon, deg, matpc):
    c = tmq.Context(X)
    c.load_gauge(tmq.gen_gauge(X), t_boundary=-1, recon=recon)
    c.set_op(KAPPA, MU, matpc)
    Vh = int(np.prod(X)) // 2
    b = c.spinor(8); b.set(tmq.gen_spinor(X, "z4")[:Vh])
    c.time_kernel(5, prec, 4, b)
    ms, per = c.time_kernel(5, prec, deg, b)
    bps = (24 * (2 + 3 + 2 + 5) + 4 * 8 * recon) * prec
    gbs = bps * Vh / (ms * 1e-3) * 1e-9
    print(json.dumps({"what": "chebyshev filter, per degree", "lattice": X, "prec": prec, "recon": recon, "matpc": matpc, "degree": deg,
                      "ms_per_degree": ms, "launches_per_degree": per, "bytes_per_site": bps, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak(),
                      "GFLOP/s": 5664.0 * Vh / (ms * 1e-3) * 1e-9}), flush=True)
    c.close()


def eig(X, nev, nkv, deg, amin, amax, tol, prec):
    c = tmq.Context(X)
    c.load_gauge(tmq.gen_gauge(X), t_boundary=-1, recon=12)
    c.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN_ASYM)
    es = c.eigset(nkv + 1, prec)
    l0 = c.launch_count()
    t0 = time.perf_counter()
    r = c.eigensolve(es, nev, nkv, poly_deg=deg, amin=amin, amax=amax, tol=tol, max_restarts=200, which=0, seed=5)
    dt = time.perf_counter() - t0
    print(json.dumps({"what": "thick-restart Lanczos", "lattice": X, "prec": prec, "nev": nev, "nkv": nkv, "degree": deg, "window": [amin, amax],
                      "tol": tol, "secs": dt, "nconv": r["nconv"], "restarts": r["restarts"], "operator_applications": r["matvecs"],
                      "launches": c.launch_count() - l0, "evals": [float(x) for x in r["evals"]],
                      "max_resid": float(r["resid"].max())}), flush=True)
    c.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96])
    ap.add_argument("--eig-lattice", type=int, nargs=4, default=[16, 16, 16, 32])
    ap.add_argument("--skip-eig", action="store_true")
    a = ap.parse_args()
    X = tuple(a.lattice)
    cheb(X, 8, 12, 20, 0)
    cheb(X, 8, 12, 20, 2)
    cheb(X, 4, 12, 20, 0)
    if not a.skip_eig:
        eig(tuple(a.eig_lattice), 16, 48, 40, 0.36, 2.0, 1e-10, 8)
