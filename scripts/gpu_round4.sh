#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary4.txt
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary4.txt
tail -n 8 gpurun_out/pytest_gpu.log
for L2 in 0 64 32; do
  DET_L2_FETCH=$L2 timeout 600 python scripts/microbench.py --tag l2fetch$L2 --dims 64 --ops find,insert,lookup_sparse,adagrad,accum > gpurun_out/microbench_l2_$L2.jsonl 2>> gpurun_out/microbench.err
  echo "microbench l2=$L2 exit $?" | tee -a gpurun_out/summary4.txt
done
timeout 900 python scripts/microbench.py --tag v2 --dims 16,128 > gpurun_out/microbench_v2_16_128.jsonl 2>> gpurun_out/microbench.err
timeout 600 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err
echo "bench exit $?" | tee -a gpurun_out/summary4.txt
cat gpurun_out/bench4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:find_kernel -s 3 -c 1 -o gpurun_out/prof_find_l2 \
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_find.log 2>&1
echo "ncu find exit $?" | tee -a gpurun_out/summary4.txt
