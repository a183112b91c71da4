#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary3.txt
for f in tests/test_table_gpu.py tests/test_fused_gpu.py tests/test_scale_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary3.txt
  tail -n 6 gpurun_out/$n.log
done
DET_FIND_VARIANT=0 DET_INSERT_VARIANT=0 timeout 600 python -m pytest tests/test_table_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 600 -k "golden or random or known" > gpurun_out/test_table_gpu_v0.log 2>&1
echo "table tests variant0 exit $?" | tee -a gpurun_out/summary3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary3.txt
timeout 1200 python scripts/microbench.py --tag v1 > gpurun_out/microbench_v1.jsonl 2> gpurun_out/microbench.err
echo "microbench v1 exit $?" | tee -a gpurun_out/summary3.txt
DET_FIND_VARIANT=0 DET_INSERT_VARIANT=0 timeout 600 python scripts/microbench.py --tag v1_ldgkeys --dims 64 --ops find,insert > gpurun_out/microbench_v1_ldg.jsonl 2>> gpurun_out/microbench.err
echo "microbench ldg exit $?" | tee -a gpurun_out/summary3.txt
timeout 600 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err
echo "bench exit $?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/bench3.json
timeout 600 python bench.py --impl reference --steps 9 --warmup 3 > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err
echo "bench ref exit $?" | tee -a gpurun_out/summary3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --resident 20000000 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/summary3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:find_kernel -s 3 -c 1 -o gpurun_out/prof_find \
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_find.log 2>&1
echo "ncu find exit $?" | tee -a gpurun_out/summary3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 99 -c 1 -o gpurun_out/prof_insert \
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_insert.log 2>&1
echo "ncu insert exit $?" | tee -a gpurun_out/summary3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:apply_kernel|segment_sum_kernel|resolve_slots|accum_kernel|remove_kernel" -s 10 -c 12 -o gpurun_out/prof_fused \
  python scripts/microbench.py --dims 64 --reps 2 --resident 20000000 --ops adagrad,lookup_sparse,accum,insert_new,remove > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused exit $?" | tee -a gpurun_out/summary3.txt
ls -la gpurun_out | head -40
