#!/bin/bash
# Round 2, GPU call 4 (gpurun --gpus 2): how fast can SM stores / TMA bulk stores write to a peer over NVLink on this box
# (scripts/nvlink_probe.py), and where does a push-exchange step spend its time (DET_XCHG_TIMING=1: CUDA events between
# the kernels of det_peer_xchg_find / _insert, printed per rank at exit).
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c4
mkdir -p $O
timeout 400 python scripts/nvlink_probe.py > $O/nvlink_probe.jsonl 2> $O/nvlink_probe.err
echo "probe exit $?"; cat $O/nvlink_probe.jsonl | cut -c1-260; tail -n 3 $O/nvlink_probe.err
DET_XCHG_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus 2 --steps 300 --warmup 10 --exchange push --no-e2e > $O/bench_push_timing.json 2> $O/bench_push_timing.err
echo "bench exit $?"; grep "det xchg timing" $O/bench_push_timing.err; tail -n 1 $O/bench_push_timing.json | cut -c1-300
