"""Multi-GPU check of the C++ QKXTM host layer (one process per rank, include/qudaQKXTM_tmq.h: initCommsGridQuda): the reference-shaped
driver qkxtm_invert_test is started by torchrun --no-python on a T-split and on a Z-split process grid and its results are compared
with the single-GPU run of the same global lattice (which tests/test_gpu_host_shim.py and tests/test_gpu_contract.py check against
the CPU oracle).  Needs 2 GPUs:   python tests/sharded_shim.py     (exit code 0 = green)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lattice_util as lu  # noqa: E402

DRV = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "qkxtm_invert_test")
G = (4, 6, 4, 8)


def run(n, local, grid, out, *flags, port=29531):
    base = [DRV, "--dim"] + [str(v) for v in local] + ["--gridsize"] + [str(v) for v in grid] + list(flags) + ["--out", out]
    cmd = base if n == 1 else [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", str(n),
                               "--master-addr", "127.0.0.1", "--master-port", str(port)] + base
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if p.returncode != 0:
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit("driver failed: " + " ".join(cmd))
    return p.stdout


def main():
    fails = []
    tmp = tempfile.mkdtemp()
    twop = ["--test", "twop", "--tol", "1e-11", "--recon", "12", "--Q_sq", "1", "--src", "1", "3", "2", "5", "--tsink", "5", "--proj", "1",
            "--nsmearGauss", "2", "--alphaGauss", "4.0"]
    ref_x = os.path.join(tmp, "ref.bin")
    run(1, G, (1, 1, 1, 1), ref_x, "--test", "invert", "--tol", "1e-11", "--recon", "12", "--prec-sloppy", "single")
    xg = np.fromfile(ref_x).reshape(-1, 4, 3, 2)                                   # global, even-odd order
    ref_t = os.path.join(tmp, "ref_tw")
    run(1, G, (1, 1, 1, 1), ref_t, *twop)
    names = ["mesons.SS.01.03.02.05.dat", "baryons.SS.01.03.02.05.dat",
             "threep_tsink5_projG5G123.proton.up.ultra_local.SS.01.03.02.05.dat", "threep_tsink5_projG5G123.proton.down.ultra_local.SS.01.03.02.05.dat"]
    for tag, grid, port in (("T", (1, 1, 1, 2), 29531), ("Z", (1, 1, 2, 1), 29535)):
        local = tuple(G[d] // grid[d] for d in range(4))
        out = os.path.join(tmp, "sh_%s.bin" % tag)
        log = run(2, local, grid, out, "--test", "invert", "--tol", "1e-11", "--recon", "12", "--prec-sloppy", "single", port=port)
        assert log.count("RESULT") == 1
        for rank in range(2):
            coord = lu.rank_coord(rank, grid)
            xl = np.fromfile(out + ".rank%d" % rank).reshape(-1, 4, 3, 2)
            e = lu.rel_l2(xl, lu.local_from_global_eo(xg, local, grid, coord))
            if not e < 1e-9:
                fails.append(("invertQuda", tag, rank, e))
        out = os.path.join(tmp, "tw_%s" % tag)
        run(2, local, grid, out, *twop, port=port + 1)
        for nm in names:
            a, b = np.loadtxt(ref_t + "." + nm), np.loadtxt(out + "." + nm)
            ncol = 4 if "mesons" in nm or "baryons" in nm else 2
            va, vb = a[:, -ncol:], b[:, -ncol:]
            e = np.abs(va - vb).max() / np.abs(va).max()
            if a.shape != b.shape or not e < 2e-5:
                fails.append((nm, tag, e))
        # MG_bench on the split lattice: the plaquette of the "smeared" links goes through the containers' ghost zones (ghostToHost ->
        # cpuExchangeGhost -> ghostToDevice, one device-side exchange) and must print the single-rank value
        import re
        want = float(re.search(r"Calculated plaquette in double precision is (\S+)", run(1, G, (1, 1, 1, 1), os.path.join(tmp, "mg1.bin"), "--test", "mgbench", "--tol", "1e-5", "--recon", "12")).group(1))
        got = float(re.search(r"Calculated plaquette in double precision is (\S+)", run(2, local, grid, os.path.join(tmp, "mg2.bin"), "--test", "mgbench", "--tol", "1e-5", "--recon", "12", port=port + 2)).group(1))
        if not abs(want - got) < 2e-6:
            fails.append(("plaquette through the ghost zones", tag, want, got))
        print("grid %s: invertQuda slabs, two-/three-point files and the container plaquette compared" % (grid,), flush=True)
    print("failures:", fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
