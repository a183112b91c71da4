#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "bench exit $?"
cat gpurun_out/bench7.json; tail -n 3 gpurun_out/bench7.err
timeout 400 python scripts/microbench.py --tag v5 --dims 64 --ops adagrad,adam,lookup_sparse > gpurun_out/microbench_v5.jsonl 2>> gpurun_out/microbench.err
cut -c 1-160 gpurun_out/microbench_v5.jsonl
timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 2>/dev/null | cut -c 1-300
