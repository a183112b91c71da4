#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary5.txt
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" | tee -a gpurun_out/summary5.txt
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary5.txt
for RU in 1 2; do
  DET_APPLY_RU=$RU timeout 600 python scripts/microbench.py --tag applyru$RU --dims 16,64,128 --ops adagrad,adam > gpurun_out/microbench_ru$RU.jsonl 2>> gpurun_out/microbench.err
  echo "microbench ru=$RU exit $?" | tee -a gpurun_out/summary5.txt
done
timeout 900 python scripts/microbench.py --tag v3 > gpurun_out/microbench_v3.jsonl 2>> gpurun_out/microbench.err
echo "microbench v3 exit $?" | tee -a gpurun_out/summary5.txt
timeout 600 python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err
echo "bench exit $?" | tee -a gpurun_out/summary5.txt
cat gpurun_out/bench5.json
timeout 600 python bench.py --workload c3 --steps 50 --warmup 5 > gpurun_out/bench5_c3.json 2> gpurun_out/bench5_c3.err
echo "bench c3 exit $?" | tee -a gpurun_out/summary5.txt
cat gpurun_out/bench5_c3.json; tail -n 3 gpurun_out/bench5_c3.err
# launch list of OUR kernels only (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"find_kernel|insert_kernel" -s 20 -c 60 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --resident 20000000 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/summary5.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|accum_kernel|remove_kernel|segment_sum|resolve_slots|export_write|export_count" -c 14 -o gpurun_out/prof_fused2 \
  python scripts/microbench.py --dims 64 --reps 1 --resident 20000000 --ops adagrad,adam,lookup_sparse_4ids,accum,insert_new,remove > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused exit $?" | tee -a gpurun_out/summary5.txt
ls gpurun_out | head -50
