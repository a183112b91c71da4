#!/bin/bash
# usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary_multi$N.txt
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_fused_gpu.py tests/test_table_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 500 > gpurun_out/test_single_on_multi.log 2>&1
  echo "single-gpu tests exit $?" | tee -a gpurun_out/summary_multi$N.txt
  tail -n 5 gpurun_out/test_single_on_multi.log
  timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 500 > gpurun_out/test_multigpu.log 2>&1
  echo "multigpu test exit $?" | tee -a gpurun_out/summary_multi$N.txt
  tail -n 15 gpurun_out/test_multigpu.log
fi
for EX in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 300 --warmup 10 --exchange $EX > gpurun_out/bench_n${N}_$EX.json 2> gpurun_out/bench_n${N}_$EX.err
  echo "bench N=$N $EX exit $?" | tee -a gpurun_out/summary_multi$N.txt
  cat gpurun_out/bench_n${N}_$EX.json | tail -n 1 | cut -c 1-1500
  tail -n 3 gpurun_out/bench_n${N}_$EX.err
done
