#!/bin/bash
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
  bench.py --gpus $N --steps 300 --warmup 10 --exchange peer > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err
echo "bench N=$N peer exit $?" | tee -a gpurun_out/summary_multi$N.txt
tail -n 1 gpurun_out/bench_n${N}_peer.json | cut -c 1-2500
grep -v "^\*\*\|OMP_NUM" gpurun_out/bench_n${N}_peer.err | tail -n 5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 \
  bench.py --gpus $N --steps 60 --warmup 5 --exchange nccl --no-e2e > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
echo "bench N=$N nccl exit $?" | tee -a gpurun_out/summary_multi$N.txt
tail -n 1 gpurun_out/bench_n${N}_nccl.json | cut -c 1-600
