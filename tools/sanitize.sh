#!/bin/bash
# compute-sanitizer over the hot path (SURVEY.md section 5 prescribes it for 8^3 x 16): memcheck, racecheck, synccheck, initcheck on
# tools/sanitize_target.py (plain, and the ghost-zone path forced in every halo mode), and memcheck on the smoke test.
# Run on the GPU box:  /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2'
# Summaries land in gpurun_out/sanitize_<tag>_*.txt; copy what should be judged into profiles/.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
summ() { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Hazard|Invalid|Uninitialized|sanitize target|Error" "$1" | sort | uniq -c | head -40; }
for tool in memcheck racecheck synccheck initcheck; do
  targets="plain mode3 mode2 mode1 mode0"
  if [ $tool = synccheck ] || [ $tool = initcheck ]; then targets="plain mode3"; fi
  for what in $targets; do
    f=$OUT/sanitize_${TAG}_${tool}_${what}.txt
    timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_target.py $what 3 > $f 2>&1
    echo "== $tool $what rc=$?" | tee -a $OUT/sanitize_${TAG}_summary.txt
    summ $f | tee -a $OUT/sanitize_${TAG}_summary.txt
  done
done
f=$OUT/sanitize_${TAG}_memcheck_smoke.txt
timeout 900 $CS --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $f 2>&1
echo "== memcheck smoke rc=$?" | tee -a $OUT/sanitize_${TAG}_summary.txt
summ $f | tee -a $OUT/sanitize_${TAG}_summary.txt
