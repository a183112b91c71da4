#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
DET_APPLY_STAGED=1 timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_peer_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 500 -k "adagrad or adam or slot_state or train_step or apply" > gpurun_out/pytest_staged.log 2>&1
echo "pytest staged exit $?"; tail -n 5 gpurun_out/pytest_staged.log
for ST in 0 1; do
  DET_APPLY_STAGED=$ST timeout 400 python scripts/microbench.py --tag staged$ST --dims 16,64,128 --ops adagrad,adam --reps 8 > gpurun_out/microbench_staged$ST.jsonl 2>> gpurun_out/microbench.err
  cut -c 1-150 gpurun_out/microbench_staged$ST.jsonl | grep uniform
done
