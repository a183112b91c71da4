#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for f in tests/test_table_gpu.py tests/test_scale_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary2.txt
  tail -n 8 gpurun_out/$n.log
done
timeout 1200 python scripts/microbench.py --tag v0 > gpurun_out/microbench_v0.jsonl 2> gpurun_out/microbench.err
echo "microbench exit $?" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench exit $?" | tee -a gpurun_out/summary2.txt
cat gpurun_out/bench2.json
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --resident 20000000 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:find_kernel -s 3 -c 2 -o gpurun_out/prof_find \
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_find.log 2>&1
echo "ncu find exit $?" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 103 -c 2 -o gpurun_out/prof_insert \
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_insert.log 2>&1
echo "ncu insert exit $?" | tee -a gpurun_out/summary2.txt
ls -la gpurun_out
