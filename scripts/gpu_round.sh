#!/bin/bash
# One gpurun call: GPU tests (each file under its own timeout), smoke, short bench, ncu launch list.
# Everything is logged under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.mem_get_info())" >> gpurun_out/host.txt 2>&1
for f in tests/test_table_gpu.py tests/test_fused_gpu.py tests/test_scale_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$n.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
