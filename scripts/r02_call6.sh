#!/bin/bash
# Round 2, GPU call 6 (gpurun --gpus N): the chunk-pipelined owner-side insert.  N=2: multi-process parity test first.
# Then bench.py at N: push with DET_XCHG_CHUNKS=4 (default) and 1 (no pipeline), timing on; peer for the A/B.
set -u
export DET_NO_REBUILD=1
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c6_n$N
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | tee $O/test_multigpu.log
  lap multigpu-test
fi
run() {  # tag, exchange, chunks
  DET_XCHG_TIMING=1 DET_XCHG_CHUNKS=$3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --steps ${STEPS:-300} --warmup 10 --exchange $2 --no-e2e > $O/bench_$1.json 2> $O/bench_$1.err
  echo "bench N=$N $1 exit $?"
  grep "det xchg timing rank 0" $O/bench_$1.err
  tail -n 1 $O/bench_$1.json | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read())
  print({k:d.get(k) for k in ('value','ms_per_step','find_ms','insert_ms')}, d['parity']['mismatches'], d['parity']['checked'], d['roofline_nvlink']['frac'], d['no_exchange']['value'])
except Exception as e: print('no line',e)
"
  tail -n 2 $O/bench_$1.err | cut -c1-300
  lap $1
}
run push_c4 push 4
run push_c1 push 1
[ "${SKIP_PEER:-0}" = "1" ] || run peer peer 1
