#!/bin/bash
# N GPUs: the e2e driver directly under torchrun --no-python with the solver's own per-solve report (halo mode, iteration-loop seconds)
# usage: bash tools/gpu_e2e_multi.sh <N> <nsrc>
N=${1:-2}; NSRC=${2:-6}
T=$((96 / N))
export TMQ_HALO_TIMEOUT_MS=20000 TMQ_COMM_ID_FILE=/tmp/tmq_e2e_id_$$ TMQ_COMM_NONCE=e2e-$$
python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  quda-qkxtm-multigrid-plugin_b200/lib/qkxtm_invert_test --dim 48 48 48 $T --gridsize 1 1 1 $N --test e2e --tol 1e-9 --niter 5000 --recon 12 \
  --kappa 0.12195121951219513 --mu 0.1 --nsrc $NSRC --e2e-reps 2 --seed 100 --verbosity-level summarize 2>&1 | grep -E "^CG:|RESULT|rror" | tail -30
