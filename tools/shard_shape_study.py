#!/usr/bin/env python
"""What the 8-GPU step of 48^3 x 96 costs that is NOT the interconnect: the 48^3 x 12 shard on ONE GPU, whole (no ghost zone) and with the
T dimension forced through the ghost-zone machinery against itself, in every halo mode.  The real 8-GPU step (0.772 ms, mode 2) equals the
self-exchanged one to 1 %, so the missing efficiency is local launch structure and small-volume kernel efficiency -- which this tool can
split up under ncu on a single GPU (usage: python tools/shard_shape_study.py [mode ...]; modes none p2p fused store nccl async pre25 pre75)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT)
import numpy as np, tmq, bench

X = tuple(int(v) for v in os.environ.get("SHARD", "48,48,48,12").split(","))
PART = tuple(int(v) for v in os.environ.get("PART", "0,0,0,1").split(","))
REPS = int(os.environ.get("REPS", "50"))
modes = sys.argv[1:] or ["none", "p2p", "fusedce", "fused", "store", "nccl"]
Vh = int(np.prod(X)) // 2
gauge = tmq.gen_gauge(X); src = tmq.gen_spinor(X, "gaussian")[:Vh]
for mode in modes:
    c = tmq.Context(X)
    if mode != "none":
        c.force_partition(PART)
        c.set_option(tmq.OPT_HALO_P2P, {"p2p": 2, "fused": 3, "fusedce": 4, "store": 1, "nccl": 0}.get(mode, 2))
        if mode == "async": c.set_option(5, 1)
        if mode.startswith("pre"): c.set_option(3, int(mode[3:]))
        if os.environ.get("PRE"): c.set_option(3, int(os.environ["PRE"]))
        if os.environ.get("DBG"): c.set_option(99, int(os.environ["DBG"]))
    c.load_gauge(gauge, recon=12); c.set_op(bench.KAPPA, bench.MU, 0)
    b = c.spinor(8); b.set(src)
    row = {"X": X, "part": PART, "mode": mode, "pre": os.environ.get("PRE", "50"), "dbg": os.environ.get("DBG", "0")}
    for kind, name in ((0, "K1"), (1, "K2"), (2, "K3"), (3, "MdagM"), (4, "step")):
        best = 1e9
        for rep in range(3):
            ms, nl = c.time_kernel(kind, 8, REPS, b)
            best = min(best, ms)
        row[name + "_ms"] = round(best, 4); row[name + "_launches"] = nl
    x = c.spinor(8)
    if not os.environ.get("DBG"):
        info = c.cg_mdagm(x, b, tol=1e-30, maxiter=200)
        row["solver_ms_per_iter"] = round(1e3 * info["loop_secs"] / info["iter"], 4)
    print(json.dumps(row), flush=True)
    c.close()
