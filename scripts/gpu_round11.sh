#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
for ST in 0 1; do
  DET_ACCUM_STAGED=$ST timeout 300 python scripts/microbench.py --tag accst$ST --dims 16,64,128 --ops accum --reps 8 > gpurun_out/microbench_accum$ST.jsonl 2>> gpurun_out/microbench.err
  cut -c 1-140 gpurun_out/microbench_accum$ST.jsonl | grep uniform
done
