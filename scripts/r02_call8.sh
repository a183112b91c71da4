#!/bin/bash
# Round 2, GPU call 8 (gpurun --gpus N): serve_find with sources dealt round-robin, the hybrid exchange (push lookups +
# one-sided inserts), the owner-side sharded optimizer step.  N=2: multi-process parity test first.
set -u
export DET_NO_REBUILD=1
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c8_n$N
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | tee $O/test_multigpu.log
  lap multigpu-test
fi
run() {  # tag, extra args...
  tag=$1; shift
  DET_XCHG_TIMING=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29615 \
    bench.py --gpus $N --warmup 10 --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "bench N=$N $tag exit $?"
  grep "det xchg timing rank 0" $O/bench_$tag.err
  tail -n 1 $O/bench_$tag.json | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read())
  print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','find_ms','insert_ms')}, (d.get('parity') or {}).get('mismatches'), (d.get('parity') or {}).get('checked'), (d.get('roofline_nvlink') or {}).get('frac'))
except Exception as e: print('no line',e)
"
  tail -n 2 $O/bench_$tag.err | cut -c1-300
  lap $tag
}
run hybrid --steps 200 --exchange hybrid
run push --steps 200 --exchange push
[ "${SKIP_PEER:-0}" = "1" ] || run peer --steps 200 --exchange peer
run c5_owner --steps 30 --workload c5 --exchange push
[ "${SKIP_PEER:-0}" = "1" ] || run c5_round1 --steps 30 --workload c5 --exchange peer
