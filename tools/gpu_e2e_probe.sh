#!/bin/bash
# e2e probe: the plug-in driver alone at 48^3x96 with per-solve timings printed (verbosity summarize)
OUT=gpurun_out
D=quda-qkxtm-multigrid-plugin_b200/lib/qkxtm_invert_test
$D --dim 48 48 48 96 --test e2e --tol 1e-9 --niter 5000 --recon 12 --nsrc ${1:-6} --e2e-reps 2 --seed 100 --verbosity-level summarize > $OUT/e2e_probe.txt 2>&1
grep -E "CG:|RESULT" $OUT/e2e_probe.txt | tail -40
