#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
echo "bench ref exit $?"; cut -c 1-400 gpurun_out/bench_final_ref.json
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"; cat gpurun_out/bench_final.json; tail -n 3 gpurun_out/bench_final.err
# launch list of the SAME bench command (profiles/rNN_launches_bench_n1.csv): per-launch times under ncu are cold-cache
# and serialised (only the table kernels are listed: the prefill's inserts come first, the LAST rows are the timed steps), so only each kernel's SHARE of the step is compared with the live CUDA-event split (find_ms / insert_ms)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:find_kernel|insert_kernel" -c 600 --csv --log-file gpurun_out/launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list exit $?"; grep -c "kernel" gpurun_out/launches_bench_n1.csv
