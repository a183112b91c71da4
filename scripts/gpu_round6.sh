#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary6.txt
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 --durations=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary6.txt
tail -n 22 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err
echo "bench exit $?" | tee -a gpurun_out/summary6.txt
cat gpurun_out/bench6.json; tail -n 3 gpurun_out/bench6.err
timeout 600 python scripts/microbench.py --tag v4 --dims 64 --ops adagrad,adam,lookup_sparse > gpurun_out/microbench_v4.jsonl 2>> gpurun_out/microbench.err
cat gpurun_out/microbench_v4.jsonl | cut -c 1-200
