#!/usr/bin/env python
"""Summarise the ncu artefacts gpurun brought back into profiles/ (tracked).
usage: tools/ncu_summary.py <tag>   reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep"""
import collections, csv, io, re, subprocess, sys

tag = sys.argv[1]
out = open("profiles/%s_summary.md" % tag, "w")
def P(*a):
    print(*a, file=out)

P("# ncu summary %s\n" % tag)
P("Command profiled: `%s` (48^3x96, 1 GPU)." % (sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu"))
P("Per-launch times below are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n")
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); ui = H.index("Metric Unit"); gi = H.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:100]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
P("## launch list (gpu__time_duration.sum)\n")
P("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    P("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))
P("\nEPI codes: 0 plain hop, 1 hop+A^-1 (K1/K3 of the CG iteration), 2 hop+A^-1+xpay, 5 = K2 (M p, fused |Mp|^2, A^-dag), "
  "7 = K4 (A^dag w - k^2 D^dag u, fused r -= alpha z and |r|^2), 6 = K2 of the asymmetric operator (A x - k^2 D t), 8 = K4 with the Chebyshev three-term recurrence fused in (EPI_CHEB).\n")
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) > 2:
    Hh, U = rr[0], rr[1]; idx = {h: i for i, h in enumerate(Hh)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    P("## ncu --set full, Dslash kernels (one row per captured launch)\n")
    for r in rr[2:]:
        P("### `%s`\n" % r[idx["Kernel Name"]][:120])
        P("| metric | value | unit |\n|---|---:|---|")
        for w in want:
            if w in idx:
                P("| %s | %s | %s |" % (w, r[idx[w]], U[idx[w]]))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
            ur, uw = U[idx["dram__bytes_read.sum"]], U[idx["dram__bytes_write.sum"]]
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tb = rd * sc[ur] + wr * sc[uw]
            P("| **traffic (read+write)** | %.4g | byte |" % tb)
        except Exception:
            pass
        P("")
out.close()
print(open("profiles/%s_summary.md" % tag).read()[:3000])
