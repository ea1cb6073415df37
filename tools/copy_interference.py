#!/usr/bin/env python
"""Does a bulk host <-> device copy slow the CG iteration of an 8-GPU shard, and through what?  ONE GPU, the 48^3 x 12 shard: unsharded
(no flags, no system-scope operation in the iteration) and with T forced through the ghost-zone machinery against itself in the halo
modes 4 / 3 / 2, timed (tmq_time_kernel kind 4) while another thread keeps the pinned-copy streams busy (tmq_host_link_probe: uploads,
then downloads, then both).  If the unsharded iteration slows as well, the cause is the link / submission path; if only the sharded
ones do, it is their system-scope flag and fence traffic.  (profiles/r2_e2e_n8.md)"""
import ctypes as C, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")); sys.path.insert(0, ROOT)
import numpy as np, tmq, bench

X = tuple(int(v) for v in os.environ.get("SHARD", "48,48,48,12").split(","))
REPS = int(os.environ.get("REPS", "40")); COPIES = int(os.environ.get("COPIES", "60"))
Vh = int(np.prod(X)) // 2
nbytes = 2 * Vh * 24 * 8
gauge = tmq.gen_gauge(X); src = tmq.gen_spinor(X, "gaussian")[:Vh]
for mode in sys.argv[1:] or ["none", "fusedce", "fused", "p2p"]:
    c = tmq.Context(X)
    if mode != "none":
        c.force_partition((0, 0, 0, 1))
        c.set_option(tmq.OPT_HALO_P2P, {"p2p": 2, "fused": 3, "fusedce": 4}[mode])
    c.load_gauge(gauge, recon=12); c.set_op(bench.KAPPA, bench.MU, 0)
    b = c.spinor(8); b.set(src)
    L = c.L
    L.tmq_host_alloc_pinned.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]
    L.tmq_host_link_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.tmq_host_free_pinned.argtypes = [C.c_void_p, C.c_void_p]
    hs, hd = C.c_void_p(), C.c_void_p()
    assert L.tmq_host_alloc_pinned(c.h, C.byref(hs), nbytes) == 0 and L.tmq_host_alloc_pinned(c.h, C.byref(hd), nbytes) == 0
    C.memset(hs, 0, nbytes)
    quiet = min(c.time_kernel(4, 8, REPS, b)[0] for _ in range(3))
    secs = (C.c_double * 3)()
    t_start = [0.0]
    def copier():
        t_start[0] = time.perf_counter()
        L.tmq_host_link_probe(c.h, hs, hd, COPIES, secs)
    th = threading.Thread(target=copier); th.start()
    time.sleep(0.01)
    samples = []
    while th.is_alive():
        t0 = time.perf_counter(); ms, _ = c.time_kernel(4, 8, REPS, b); t1 = time.perf_counter()
        samples.append((t0 - t_start[0], t1 - t_start[0], ms))
    th.join()
    edges = [0.0, secs[0], secs[0] + secs[1], secs[0] + secs[1] + secs[2]]
    row = {"X": X, "mode": mode, "quiet_ms": round(quiet, 4), "copy_GBps": [round(nbytes * COPIES * (2 if i == 2 else 1) / secs[i] * 1e-9, 1) for i in range(3)]}
    for i, name in enumerate(("during_uploads_ms", "during_downloads_ms", "during_both_ms")):
        inside = [ms for (a, e, ms) in samples if a >= edges[i] and e <= edges[i + 1]]
        row[name] = round(float(np.median(inside)), 4) if inside else None
        row[name.replace("_ms", "_n")] = len(inside)
    print(json.dumps(row), flush=True)
    L.tmq_host_free_pinned(c.h, hs); L.tmq_host_free_pinned(c.h, hd)
    c.close()
