#!/usr/bin/env python
"""bench.py -- the even-odd twisted-mass Dslash inside CG on M^dag M, 48^3x96 fp64 (BASELINE.json metric).

One JSON line on stdout (rank 0).  Definitions (DESIGN.md section "Measurement"):
  step     one CG iteration on M_pc^dag M_pc = 4 even-odd TM Dslash launches (fused twists / xpay /
           reductions) + 1 fused update launch, on a resident 48^3x96 (global) lattice, no host sync
  value    GFLOP/s over the K timed steps, QUDA flop model 5904 flop per parity site per iteration
           (SURVEY.md 8d), CUDA events on the launching stream, max over ranks
  solver_loop  the same iteration inside the real solver (tmq_cg_mdagm, stopping test included)
  e2e      the same metric through the reference-facing plug-in calls of libqkxtm_tmq.so with HOST buffers (one process per rank running
           the qkxtm_invert_test driver): invertMultiSrcQuda over 12 right-hand sides -- pinned host source -> H2D -> prepare -> M^dag ->
           CG to tol (real iteration count) -> reconstruct -> D2H per column, every copy inside the timed region, the copies of the
           neighbouring columns behind each solve; flops = 5904 x Vh x iterations.  The single invertQuda (fp64 and fp32-sloppy) is
           reported beside it: one solve cannot hide its own copies
  roofline the hop + A^-1 Dslash kernel (EPI_TW, half of all Dslash launches) timed alone with CUDA events:
           algorithmic bytes (24 + 24 + 8*12) * 8 = 1152 B per output parity site (fp64, recon 12)
  scale64  BASELINE.json configs[4], 64^3x128 sharded T then Z (T x Z at 8 GPUs): ms per fused CG iteration and the checksum of a solve
  f_rows   N = 1: one Chebyshev degree, the twisted-clover hop + A^-1 and CG iteration, one Gaussian smearing step, each
           timed with CUDA events inside libtmq and set against its algorithmic bytes
  e2e.host_link  pinned host <-> device bandwidth with every rank of the job copying at once (what the box gives the e2e legs)
  cpu_baseline / --impl reference: the CPU oracle (oracle/, a port: the reference's own host code cannot be
           built here) running CG iterations of the same 48^3x96 workload on all host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

KAPPA = 1.0 / (2.0 * (4.0 + 0.1))
MU = 0.1
FLOPS_ITER = 5904.0           # per parity site per CG iteration (M^dag M 5664 + blas 240)
FLOPS_K = {0: 1320.0, 1: 1392.0, 2: 1440.0}
KNAME = {0: "K1 hop", 1: "K2 hop+A^-1", 2: "K3 hop+A^-1+xpay"}


def bytes_per_site(kind, prec, recon):
    spinors = {0: 2, 1: 2, 2: 3}[kind]
    return (24 * spinors + 8 * recon) * prec


def step_bytes_per_site(prec, recon):
    # K1 (in,out) + K2 (in,x,out) + K3 (in,out) + K4 (in,x,r read,r write) + update (x,p r/w, r read) + 4 gauge sweeps
    return (24 * (2 + 3 + 2 + 4 + 5) + 4 * 8 * recon) * prec


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples drawing more than half of the maximum observed power
        pmax = max(power)
        load = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "power_w_max": pmax,
                "samples": len(sm), "reasons": sorted(reasons)}


def choose_grid(n):
    """T first, then Z (BASELINE.json: 'sharded along T, then Z')."""
    return {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 1, 4), 8: (1, 1, 1, 8)}[n]


def dist_setup(n):
    if n == 1:
        return None, 0
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    assert dist.get_world_size() == n, "launch with torchrun --nproc-per-node %d" % n
    return dist, dist.get_rank()


def dist_max(dist, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([float(v)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def dist_sum(dist, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([float(v)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


# ---------------------------------------------------------------------------------------------------------------------
def workload_config(GX):
    """`config` of BOTH arms (the driver compares them): the workload BASELINE.json's metric is quoted on"""
    return {"workload": "%dx%dx%dx%d even-odd twisted-mass Dslash in CG on MdagM, fp64" % tuple(GX), "lattice": list(GX),
            "kappa": KAPPA, "mu": MU, "matpc": "even-even", "gauge": "random SU(3), anti-periodic T, seed 137", "source": "Z4 noise, seed 100"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle(X):
    """the CPU oracle with ALL host cores: launchers such as torchrun export OMP_NUM_THREADS=1 to every worker, so the thread
    count is set explicitly"""
    from oracle.oracle import Oracle
    o = Oracle(X)
    o.set_num_threads(host_cores())
    return o


def numpy_fields(X):
    """synthetic gauge field (QDP even-odd, anti-periodic T folded in) and the even half of a Z4 source, generated with numpy only
    (tests/lattice_util.py: the same counter-based generator as the product's host library, which the CPU arms must not load), in
    blocks of time slices to bound the memory"""
    import lattice_util as lu
    V = int(np.prod(X)); Vh = V // 2
    tc = 8 if X[3] % 8 == 0 else 2
    nch = X[3] // tc
    Xc = (X[0], X[1], X[2], tc)
    Vch = int(np.prod(Xc)) // 2
    grid = (1, 1, 1, nch)
    gauge = np.empty((4, V, 3, 3, 2), dtype=np.float64)
    rhs = np.empty((Vh, 4, 3, 2), dtype=np.float64)
    def block(c):
        g = lu.random_gauge_qdp(Xc, 137, -1, grid, (0, 0, 0, c))
        gauge[:, c * Vch:(c + 1) * Vch] = g[:, :Vch]
        gauge[:, Vh + c * Vch: Vh + (c + 1) * Vch] = g[:, Vch:]
        s = lu.spinor_eo_from_lex(lu.z4_source_lex(Xc, 100, grid, (0, 0, 0, c)), Xc)
        rhs[c * Vch:(c + 1) * Vch] = s[:Vch]
    from concurrent.futures import ThreadPoolExecutor      # numpy releases the GIL in its array loops
    with ThreadPoolExecutor(max_workers=min(host_cores(), 32)) as ex:
        list(ex.map(block, range(nch)))
    return gauge, rhs


def cpu_cg_sample(X, budget_s, gauge, rhs, iters=None):
    """the CPU oracle (port) timed on this box: `iters` CG iterations (or as many as fit in budget_s)."""
    o = cpu_oracle(X)
    t0 = time.perf_counter()
    o.mdagm(gauge, rhs, KAPPA, MU, 0)                      # calibration (and page-in)
    t1 = time.perf_counter() - t0
    if iters is None:
        iters = int(max(2, min(50, budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    _, it, _, _ = o.cg_mdagm(gauge, rhs, KAPPA, MU, 0, tol=1e-30, maxiter=iters)
    dt = time.perf_counter() - t0
    gf = FLOPS_ITER * o.Vh * it / dt * 1e-9
    return {"value": gf, "unit": "GFLOP/s", "cores": o.num_threads(), "kind": "port",
            "sample": "%dx%dx%dx%d fp64, %d CG iterations on MdagM (%.2f s)" % (tuple(X) + (it, dt)), "ms_per_iter": dt / it * 1e3,
            "note": "scalar C99 + OpenMP port of the host reference (-ffp-contract=off), a reported baseline and not an optimised CPU code"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Its own host Dslash/CG is upstream QUDA
    (tests/wilson_dslash_reference.cpp), absent from /root/reference and not buildable here, so the timed code
    is the oracle port with all host threads.  Each step = one CPU CG iteration on M_pc^dag M_pc of the SAME 48^3x96 workload (bounded
    by the number of steps: --steps 20 --warmup 5 take about 40 s on 16 cores).  Only the oracle library is loaded: fields come from numpy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    X = tuple(args.lattice)
    o = cpu_oracle(X)
    gauge, rhs = numpy_fields(X)
    if warm > 0:
        o.cg_mdagm(gauge, rhs, KAPPA, MU, 0, tol=1e-30, maxiter=warm)
    t0 = time.perf_counter()
    _, it, _, _ = o.cg_mdagm(gauge, rhs, KAPPA, MU, 0, tol=1e-30, maxiter=steps)
    dt = time.perf_counter() - t0
    gf = FLOPS_ITER * o.Vh * it / dt * 1e-9
    sample = "%d CG iterations on MdagM of the %dx%dx%dx%d fp64 workload per run (%d warm-up)" % ((it,) + X + (warm,))
    line = {"impl": "reference", "metric": "tm_dslash_cg_gflops", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / max(it, 1) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(X),
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": o.num_threads(), "kind": "port", "sample": sample,
                             "note": "scalar C99 + OpenMP port of the host reference (-ffp-contract=off); upstream QUDA's own host code is not in /root/reference"},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def kernel_source_hash():
    """sha256 of the Dslash kernel sources: an ncu traffic figure is only quoted for the sources it was captured on"""
    import hashlib
    h = hashlib.sha256()
    for f in ("tmq_site.cuh", "tmq_dslash_inst.cuh", "tmq_types.h"):
        with open(os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def committed_traffic(prec, recon, X):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), if one exists
    for this exact kernel, local lattice AND kernel sources (the capture records their hash); else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t.get("kernel_source_sha256_16") != kernel_source_hash():
            return None, "capture %s is of older kernel sources" % t.get("capture", "?")
        return t.get("dslash_kernel<%s,%d,EPI_TW>@%dx%dx%dx%d" % (("double" if prec == 8 else "float", recon) + tuple(X))), t.get("capture")
    except Exception:
        return None, None


def f_rows_leg(tmq, ctx, b_par, prec, recon, X, peak):
    """SURVEY 8f rows on the resident lattice, each timed with CUDA events inside libtmq and set against its algorithmic bytes
    (DESIGN.md section 4).  Never fatal: a failure is reported in the key instead of losing the bench line."""
    out = {}
    Vh = int(np.prod(X)) // 2
    gbs = lambda bytes_per_site, sites, ms: bytes_per_site * sites / (ms * 1e-3) * 1e-9
    try:   # f1: one degree of the Chebyshev-accelerated MdagM (4 Dslash launches, recurrence fused into the 4th)
        ms, nl = ctx.time_kernel(5, prec, 10, b_par)
        bps = (12 * 24 + 32 * recon) * prec
        out["chebyshev_degree"] = {"ms": ms, "launches": nl, "bytes_per_site": bps, "GB/s": gbs(bps, Vh, ms), "frac_of_hbm_peak": gbs(bps, Vh, ms) / peak}
    except Exception as e:
        out["chebyshev_degree"] = {"error": str(e)[:200]}
    try:   # f3: twisted-clover, hop + A^-1 and the fused CG iteration (site matrices: 72 complex per site and application)
        ctx.clover_load(1.57551 * KAPPA)
        ms1, _ = ctx.time_kernel(1, prec, 10, b_par)
        ms4, _ = ctx.time_kernel(4, prec, 10, b_par)
        cl = 72 * 2 * prec
        b1 = (24 + 24 + 8 * recon) * prec + cl
        b4 = (24 * 16 + 32 * recon) * prec + 3 * cl + 24 * prec
        out["twisted_clover"] = {"hop+Ainv_ms": ms1, "hop+Ainv_bytes_per_site": b1, "hop+Ainv_frac_of_hbm_peak": gbs(b1, Vh, ms1) / peak,
                                 "cg_iter_ms": ms4, "cg_iter_bytes_per_site": b4, "cg_iter_frac_of_hbm_peak": gbs(b4, Vh, ms4) / peak}
        ctx.clover_free()
    except Exception as e:
        out["twisted_clover"] = {"error": str(e)[:200]}
        try: ctx.clover_free()
        except Exception: pass
    try:   # f2: Gaussian smearing in the containers' layout, streaming order, fp64
        V = 2 * Vh
        nb_vec, nb_g = V * 24 * 8, V * 72 * 8
        bufs = [ctx.dev_malloc(nb_vec), ctx.dev_malloc(nb_vec), ctx.dev_malloc(nb_g)]
        try:
            for pp, nn in zip(bufs, (nb_vec, nb_vec, nb_g)):
                ctx.L.tmq_dev_memset(ctx.h, pp, 0x3c, nn)           # small normal numbers (timing only; parity is tests/test_gpu_smear.py)
            ctx.qkxtm_gauss_smear(bufs[1], bufs[0], bufs[2], 8, 2, 0.25)
            ctx.timer_start()
            ctx.qkxtm_gauss_smear(bufs[1], bufs[0], bufs[2], 8, 20, 0.25)
            ms = ctx.timer_stop() / 20
            bps = (24 + 24 + 54) * 8
            out["gauss_smearing_step"] = {"ms": ms, "bytes_per_site": bps, "GB/s": gbs(bps, V, ms), "frac_of_hbm_peak": gbs(bps, V, ms) / peak}
        finally:
            for pp in bufs:
                ctx.dev_free(pp)
    except Exception as e:
        out["gauss_smearing_step"] = {"error": str(e)[:200]}
    return out


def make_context(tmq, dist, rank, local_rank, X, grid, coord, args, n):
    ctx = tmq.Context(X, grid=grid, coord=coord, device=local_rank)
    if n > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(tmq.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(uid, src=0)
        ctx.comm_init(bytes(uid.numpy().tobytes()), n, rank)
    if args.tile:
        ctx.set_tile(*args.tile)
    ctx.set_option(tmq.OPT_HALO_P2P, {"fusedce": 4, "fused": 3, "p2p": 2, "store": 1, "nccl": 0}[args.halo])
    if args.boundary_at is not None:
        ctx.set_option(3, args.boundary_at)
    if args.pack_async is not None:
        ctx.set_option(5, args.pack_async)
    return ctx


HALO_NAMES = {0: "none", 1: "nccl send/recv", 2: "peer-memory stores + fused launch", 3: "copy-engine peer copies + fused launch",
              4: "fused compute + halo exchange: boundary CTAs store the next application's faces into the neighbours' arenas",
              5: "producer's boundary CTAs pack the next application's faces locally, copy-engine peer copies + fused launch"}


def e2e_through_the_plugin(args, n, rank, local_rank, GX, grid, dist):
    """the e2e leg: the C++ QKXTM shim (libqkxtm_tmq.so: initQuda -> loadGaugeQuda -> invertQuda / invertMultiSrcQuda) driven by the
    qkxtm_invert_test driver, one process per rank, page-locked HOST sources and solutions, every copy inside the timed region"""
    drv = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "qkxtm_invert_test")
    X = tuple(GX[d] // grid[d] for d in range(4))
    env = dict(os.environ)
    env["TMQ_COMM_ID_FILE"] = "/tmp/tmq_bench_id_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid())
    env["TMQ_COMM_NONCE"] = "bench-%d-%s" % (os.getppid(), os.environ.get("TORCHELASTIC_RUN_ID", "-"))
    env["LOCAL_RANK"] = str(local_rank)
    env["TMQ_HALO_P2P"] = str({"fusedce": 4, "fused": 3, "p2p": 2, "store": 1, "nccl": 0}[args.halo])     # the same ghost exchange as the device-resident legs
    env.pop("OMP_NUM_THREADS", None)                     # the host-side field generator may use the cores
    cmd = [drv, "--dim"] + [str(v) for v in X] + ["--gridsize"] + [str(v) for v in grid] + \
          ["--test", "e2e", "--tol", repr(args.tol), "--niter", str(args.maxiter), "--recon", str(args.recon), "--kappa", repr(KAPPA), "--mu", repr(MU),
           "--nsrc", str(args.e2e_columns), "--e2e-reps", str(args.e2e_solves), "--seed", "100", "--verbosity-level", "silent"]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
    if p.returncode != 0:
        raise RuntimeError("e2e driver failed on rank %d: %s" % (rank, (p.stdout + p.stderr)[-2000:]))
    if rank != 0:
        return None
    import re
    r = json.loads(re.search(r"RESULT_E2E (\{.*\})", p.stdout).group(1))
    Vh_glob = int(np.prod(GX)) // 2
    field = r["bytes_per_field"] * n
    gf = lambda it, secs: FLOPS_ITER * Vh_glob * it / secs * 1e-9
    return {"value": gf(r["multi_iter"], r["multi_secs"]), "unit": "GFLOP/s",
            "h2d_bytes_per_step": int(field * r["nsrc"] / max(r["multi_iter"], 1)), "d2h_bytes_per_step": int(field * r["nsrc"] / max(r["multi_iter"], 1)),
            "what": "invertMultiSrcQuda of libqkxtm_tmq.so, %d right-hand sides (a point-to-all propagator has 12 columns), fp64, tol %g: pinned host source -> H2D -> "
                    "prepare -> Mdag -> CG -> reconstruct -> D2H per column, the copies of columns k+-1 behind the solve of column k; step = one CG "
                    "iteration, bytes amortised per iteration" % (r["nsrc"], args.tol),
            "columns": r["nsrc"], "secs": r["multi_secs"], "iterations": r["multi_iter"], "true_res": r["multi_true_res"],
            "solver_secs": r["multi_solver_secs"], "solver_ms_per_iter": r["multi_solver_secs"] / max(r["multi_iter"], 1) * 1e3,
            "halo": HALO_NAMES.get(r.get("halo_mode"), "?"), "last_column_cg_loop_secs": r.get("last_column_loop_secs"),
            "h2d_bytes_total": int(field * r["nsrc"]), "d2h_bytes_total": int(field * r["nsrc"]),
            "host_link": {"what": "pinned host <-> device copies of the same fields, all %d ranks at once, summed over the ranks (tmq_host_link_probe): "
                                  "what the box gives the e2e legs; the pipelined leg moves h2d_bytes_total + d2h_bytes_total in `secs`" % n,
                          "h2d_GBps": r.get("link_h2d_gbs"), "d2h_GBps": r.get("link_d2h_gbs"), "duplex_GBps": r.get("link_duplex_gbs"),
                          "pipelined_leg_GBps": 2 * field * r["nsrc"] / r["multi_secs"] * 1e-9},
            "single_solve": {"what": "one invertQuda, fp64 (nothing to hide the 2 x %.2f GB of PCIe traffic behind)" % (field / 1e9), "value": gf(r["single_iter"], r["single_secs"]),
                             "secs": r["single_secs"], "iterations": r["single_iter"], "true_res": r["single_true_res"]},
            "single_solve_mixed": {"what": "one invertQuda, cuda_prec_sloppy = single, reliable_delta = 1e-4 (the reference drivers' setting)",
                                   "value": gf(r["mixed_iter"], r["mixed_secs"]), "secs": r["mixed_secs"], "iterations": r["mixed_iter"], "true_res": r["mixed_true_res"]}}


def scale64_leg(tmq, args, n, rank, local_rank, dist):
    """BASELINE.json configs[4]: 64^3 x 128, T then Z (T x Z at 8 GPUs), one fused CG iteration per step + a CG solve whose solution
    checksum (sum of all components, |x|^2) must not depend on N"""
    GX = (64, 64, 64, 128)
    grid = {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 1, 4), 8: (1, 1, 2, 4)}[n]
    X = tuple(GX[d] // grid[d] for d in range(4))
    coord = (0, 0, (rank // grid[3]) % grid[2], rank % grid[3])
    Vh_loc = int(np.prod(X)) // 2
    ctx = make_context(tmq, dist, rank, local_rank, X, grid, coord, args, n)
    gauge = tmq.gen_gauge(X, seed=137, t_boundary=-1, grid=grid, coord=coord)
    ctx.load_gauge(gauge, t_boundary=-1, recon=args.recon)
    del gauge
    ctx.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    src = tmq.gen_spinor(X, "z4", seed=100, grid=grid, coord=coord)
    b = ctx.spinor(8); b.set(src[:Vh_loc])
    del src
    ctx.time_kernel(4, 8, 3, b)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, _ = ctx.time_kernel(4, 8, max(args.steps, 10), b)
    ms = dist_max(dist, ms)
    x = ctx.spinor(8)
    info = ctx.cg_mdagm(x, b, tol=1e-9, maxiter=2000)
    clocks = sampler.stop() if rank == 0 else None
    ones = ctx.spinor(8); ones.set(np.ones((Vh_loc, 4, 3, 2)))
    sum_x = ctx.redot(ones, x)                                   # all-reduced over the ranks inside libtmq
    norm2_x = ctx.norm2(x)
    out = {"lattice": list(GX), "grid": list(grid), "local_lattice": list(X), "halo": HALO_NAMES[ctx.halo_mode()], "ms_per_step": ms,
           "value": FLOPS_ITER * (int(np.prod(GX)) // 2) / (ms * 1e-3) * 1e-9, "unit": "GFLOP/s",
           "step": "1 fused CG iteration (4 Dslash + 1 update launch), fp64 recon %d" % args.recon,
           "solution_checksum": {"sum_x": sum_x, "norm2_x": norm2_x, "iterations": info["iter"], "true_res": info["true_res"], "tol": 1e-9,
                                 "solver_ms_per_iter": info["secs"] / max(info["iter"], 1) * 1e3},
           "clocks": clocks}
    ctx.close()
    return out


def run_native(args):
    import tmq
    n = args.gpus
    dist, rank = dist_setup(n)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    GX = tuple(args.lattice)
    grid = tuple(args.grid) if args.grid else choose_grid(n)
    assert int(np.prod(grid)) == n
    if args.scaling == "weak":
        X = GX
        GX = tuple(GX[d] * grid[d] for d in range(4))
    else:
        assert all(GX[d] % grid[d] == 0 for d in range(4))
        X = tuple(GX[d] // grid[d] for d in range(4))
    coord = (0, 0, (rank // grid[3]) % grid[2], rank % grid[3])
    prec, recon = args.prec, args.recon
    Vh_loc = int(np.prod(X)) // 2
    Vh_glob = int(np.prod(GX)) // 2

    ctx = make_context(tmq, dist, rank, local_rank, X, grid, coord, args, n)
    halo_mode = HALO_NAMES[ctx.halo_mode()]
    gauge = tmq.gen_gauge(X, seed=137, t_boundary=-1, grid=grid, coord=coord)
    ctx.load_gauge(gauge, t_boundary=-1, recon=recon)
    ctx.set_op(KAPPA, MU, tmq.MATPC_EVEN_EVEN)
    src_full = tmq.gen_spinor(X, "z4", seed=100, grid=grid, coord=coord)
    b_par = ctx.spinor(8); b_par.set(src_full[:Vh_loc])
    del src_full

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    # ---- value: K fused CG iterations, device resident, CUDA events inside libtmq on its own stream
    sampler = ClockSampler(local_rank)
    ctx.time_kernel(4, prec, max(args.warmup, 3), b_par)          # untimed warm-up steps
    barrier()
    l0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    ms_step, launches_per_step = ctx.time_kernel(4, prec, args.steps, b_par)
    barrier()
    wall = time.perf_counter() - wall0
    # per-kernel timings (same stream, CUDA events), used for the roofline
    kern = {}
    for kind in (0, 1, 2):
        ms, _ = ctx.time_kernel(kind, prec, max(args.steps, 10), b_par)
        kern[kind] = dist_max(dist, ms)
    # the REAL solver loop (tmq_cg_mdagm: the same kernels plus the stopping test): K iterations, wall clock inside the library
    xs = ctx.spinor(8)
    ctx.cg_mdagm(xs, b_par, tol=1e-30, maxiter=3)
    barrier()
    loop = ctx.cg_mdagm(xs, b_par, tol=1e-30, maxiter=max(args.steps, 10))
    solver_ms = dist_max(dist, loop["loop_secs"] / max(loop["iter"], 1) * 1e3)
    xs.free()
    clocks = sampler.stop() if rank == 0 else None
    l1 = ctx.launch_count()
    ms_step = dist_max(dist, ms_step)
    value = FLOPS_ITER * Vh_glob / (ms_step * 1e-3) * 1e-9

    peak, peak_src = measured_peak()
    bps = bytes_per_site(1, prec, recon)
    ach = bps * Vh_loc / (kern[1] * 1e-3) * 1e-9
    traffic, capture = (args.ncu_traffic, "command line") if args.ncu_traffic is not None else committed_traffic(prec, recon, X)
    roofline = {"bound": "hbm", "kernel": "dslash_kernel<%s,%d,EPI_TW> (hop + A^-1)" % ("double" if prec == 8 else "float", recon),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_site": bps, "sites_per_launch": Vh_loc, "ms_per_launch": kern[1],
                "traffic": traffic, "traffic_source": capture}
    kernels = {}
    for kind in (0, 1, 2):
        bb = bytes_per_site(kind, prec, recon)
        kernels[KNAME[kind]] = {"ms": kern[kind], "GB/s": bb * Vh_loc / (kern[kind] * 1e-3) * 1e-9,
                                "GFLOP/s": FLOPS_K[kind] * Vh_loc / (kern[kind] * 1e-3) * 1e-9,
                                "frac_of_hbm_peak": bb * Vh_loc / (kern[kind] * 1e-3) * 1e-9 / peak}
    step_gbs = step_bytes_per_site(prec, recon) * Vh_loc / (ms_step * 1e-3) * 1e-9

    # ---- the "next" rows of SURVEY 8f on the same resident lattice (N = 1 only; a few seconds): driver-visible numbers for the
    #      Chebyshev filter, the twisted-clover iteration and the Gaussian smearing, each against its own algorithmic bytes
    frows = None
    if n == 1 and not args.no_frows:
        frows = f_rows_leg(tmq, ctx, b_par, prec, recon, X, peak)

    # ---- CPU baseline on rank 0, N=1 only (bounded sample of the same workload)
    cpu = None
    if n == 1 and not args.no_cpu:
        rhs_np = np.ascontiguousarray(b_par.get())
        cpu = cpu_cg_sample(X, args.cpu_budget, gauge=gauge, rhs=rhs_np)
        del rhs_np
    del gauge
    ctx.close()

    # ---- e2e: through the plug-in's host-facing calls, HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = e2e_through_the_plugin(args, n, rank, local_rank, GX, grid, dist)
        if dist is not None:
            dist.barrier()

    # ---- 64^3 x 128 (BASELINE.json configs[4])
    s64 = None
    if args.scale64 and args.scaling == "strong":
        s64 = scale64_leg(tmq, args, n, rank, local_rank, dist)

    if rank == 0:
        cfg = workload_config(GX)
        cfg["l2_policy"] = "inputs larger than L2 at every N: per step one GPU streams 4 gauge sweeps + 16 spinor fields = %.1f GB at N = %d (126 MB L2)" \
                           % (step_bytes_per_site(prec, recon) * Vh_loc / 1e9, n)
        line = {"metric": "tm_dslash_cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": n, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64" if prec == 8 else "f32", "data": "synthetic",
                "config": cfg,
                "run": {"local_lattice": list(X), "grid": list(grid), "halo": halo_mode, "recon": recon,
                        "step": "1 CG iteration = 4 Dslash launches + 1 fused update launch, no host sync (tmq_time_kernel kind 4)",
                        "timing": "CUDA events on libtmq's compute stream, max over ranks; wall %.3f s" % wall},
                "solver_loop": {"what": "the same iteration inside tmq_cg_mdagm: its iteration loop with the stopping test (taken on the device; the host reads |r|^2 one iteration behind the launches), wall clock inside the library, max over ranks",
                                "ms_per_iter": solver_ms, "iterations": loop["iter"], "value": FLOPS_ITER * Vh_glob / (solver_ms * 1e-3) * 1e-9},
                "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
                "gpu_launches_total": int(l1 - l0),
                "step_hbm_gbs": step_gbs, "step_frac_of_hbm_peak": step_gbs / peak,
                "kernels": kernels, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "scale64": s64, "f_rows": frows}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--lattice", type=int, nargs=4, default=[48, 48, 48, 96], help="global lattice (strong) / local (weak)")
    ap.add_argument("--grid", type=int, nargs=4, default=None)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--prec", type=int, default=8, choices=[8, 4])
    ap.add_argument("--recon", type=int, default=12, choices=[8, 12, 18])
    ap.add_argument("--tile", type=int, nargs=3, default=None)
    ap.add_argument("--boundary-at", type=int, default=None, help="%% of interior CTAs scheduled before the boundary CTAs")
    ap.add_argument("--pack-async", type=int, default=None, help="1: launch the face pack on the exchange stream (TMQ_OPT_PACK_ASYNC)")
    ap.add_argument("--halo", default="fusedce", choices=["fusedce", "fused", "p2p", "store", "nccl"],
                    help="ghost exchange (TMQ_OPT_HALO_P2P): fusedce = the producing launch packs the next faces locally + copy-engine peer copies (default); "
                         "p2p = pack launch + copy-engine peer copies; fused / store = peer stores by the producing launch / the pack launch; nccl = ncclSend/Recv")
    ap.add_argument("--tol", type=float, default=1e-9)
    ap.add_argument("--maxiter", type=int, default=5000)
    ap.add_argument("--sloppy-prec", type=int, default=8, choices=[8, 4])
    ap.add_argument("--delta", type=float, default=1e-4, help="reliable_delta of the mixed-precision solve (the reference drivers set 1e-4, qkxtm/Calc_Loops.cpp:481)")
    ap.add_argument("--e2e-solves", type=int, default=2, help="repetitions of the single-solve legs of the e2e measurement")
    ap.add_argument("--e2e-columns", type=int, default=12, help="right-hand sides of the pipelined e2e leg (a propagator has 12 columns)")
    ap.add_argument("--no-frows", action="store_true", help="skip the Chebyshev / clover / smearing timings (N = 1 only)")
    ap.add_argument("--scale64", type=int, default=1, help="also time 64^3x128 (BASELINE.json configs[4]) and emit its solution checksum")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ncu-traffic", type=float, default=None, help="dram bytes per launch from the committed ncu capture")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
