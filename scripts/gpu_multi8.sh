#!/bin/bash
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 \
  bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err
echo "bench N=$N exit $?"
tail -n 1 gpurun_out/bench_n${N}_final.json | cut -c 1-3000
grep -v "^\*\*\|OMP_NUM" gpurun_out/bench_n${N}_final.err | tail -n 4
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 \
  bench.py --gpus $N --workload c5 --steps 30 --warmup 3 > gpurun_out/bench_n${N}_c5.json 2> gpurun_out/bench_n${N}_c5.err
echo "bench c5 N=$N exit $?"
tail -n 1 gpurun_out/bench_n${N}_c5.json | cut -c 1-1500
grep -v "^\*\*\|OMP_NUM" gpurun_out/bench_n${N}_c5.err | tail -n 6
# BASELINE configs[3]: the same Find + Insert step on a 1B-key table (125M resident keys per GPU), dim 128
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29653 \
  bench.py --gpus $N --workload c4 --steps 200 --warmup 10 --no-e2e > gpurun_out/bench_n${N}_c4.json 2> gpurun_out/bench_n${N}_c4.err
echo "bench c4 N=$N exit $?"
tail -n 1 gpurun_out/bench_n${N}_c4.json | cut -c 1-1500
grep -v "^\*\*\|OMP_NUM" gpurun_out/bench_n${N}_c4.err | tail -n 4
