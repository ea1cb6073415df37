#!/usr/bin/env python
"""Instruction mix of the hot kernels from the sm_100a SASS inside libtmq.so (no GPU needed): which memory instructions the Dslash issues
(LDG width, cache policy), how much fp64 arithmetic, resources per kernel.  usage: tools/sass_summary.py > profiles/<tag>_sass.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "libtmq.so")
WANT = [("dslash_kernel<double,12,EPI_TW> (roofline kernel)", "_ZN3tmq13dslash_kernelIdLi12ELi1ELb0ELb0EEEvNS_10DslashArgsIT_EE"),
        ("dslash_kernel<double,12,EPI_MDAGM2>", "_ZN3tmq13dslash_kernelIdLi12ELi5ELb0ELb0EEEvNS_10DslashArgsIT_EE"),
        ("dslash_kernel<double,12,EPI_CG4>", "_ZN3tmq13dslash_kernelIdLi12ELi7ELb0ELb0EEEvNS_10DslashArgsIT_EE"),
        ("dslash_kernel<double,12,EPI_TW>, sharded (MULTI: ghost zones, flag waits, fused pack)", "_ZN3tmq13dslash_kernelIdLi12ELi1ELb1ELb0EEEvNS_10DslashArgsIT_EE"),
        ("dslash_kernel<float,12,EPI_TW>", "_ZN3tmq13dslash_kernelIfLi12ELi1ELb0ELb0EEEvNS_10DslashArgsIT_EE")]
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
lines = res.splitlines()
for i, l in enumerate(lines):
    m = re.match(r"\s*Function (\S+):", l)
    if m and i + 1 < len(lines):
        usage[m.group(1)] = lines[i + 1].strip()
print("# SASS summary of the hot kernels (sm_100a cubins inside libtmq.so; `tools/sass_summary.py`)\n")
for title, sym in WANT:
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for l in sass.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1)] += 1
    if not ops:
        print("## %s\n\nnot found (%s)\n" % (title, sym)); continue
    tot = sum(ops.values())
    grp = lambda pred: sum(v for k, v in ops.items() if pred(k))
    print("## %s\n" % title)
    print("`%s`\n" % usage.get(sym, "?"))
    print("| | count |\n|---|---:|")
    print("| instructions | %d |" % tot)
    print("| fp64 arithmetic (DFMA / DADD / DMUL) | %d / %d / %d |" % (grp(lambda k: k.startswith("DFMA")), grp(lambda k: k.startswith("DADD")), grp(lambda k: k.startswith("DMUL"))))
    print("| fp32 arithmetic (FFMA / FADD / FMUL) | %d / %d / %d |" % (grp(lambda k: k.startswith("FFMA")), grp(lambda k: k.startswith("FADD")), grp(lambda k: k.startswith("FMUL"))))
    for pre in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "MEMBAR", "BAR", "SHFL"):
        sub = {k: v for k, v in ops.items() if k.startswith(pre)}
        if sub:
            print("| %s | %s |" % (pre, ", ".join("`%s` x %d" % kv for kv in sorted(sub.items(), key=lambda kv: -kv[1]))))
    print("| tensor-core / TMA (`UTC*`, `UTMA*`, `HMMA`, `DMMA`) | %d (an HBM-bound stencil: none expected) |" % grp(lambda k: k.startswith(("UTC", "UTMA", "HMMA", "DMMA", "QMMA"))))
    print()
