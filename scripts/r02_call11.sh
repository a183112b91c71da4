#!/bin/bash
# Round 2, GPU call 11 (gpurun --gpus 2): the EXACT command the driver runs at N=2 (default exchange, e2e included), then
# the register-capped one-sided insert kernel (DET_PEER_MINB=4) as an A/B of the hybrid step.
set -u
export DET_NO_REBUILD=1
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c11_n$N
mkdir -p $O
run() {  # tag, env assignments..., -- bench args
  tag=$1; shift
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29616 \
    bench.py --gpus $N --steps 200 --warmup 10 $EXTRA > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "bench N=$N $tag exit $?"
  tail -n 1 $O/bench_$tag.json | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read())
  print({k:d.get(k) for k in ('value','ms_per_step','find_ms','insert_ms')}, d['parity']['mismatches'], d['parity']['checked'], (d.get('e2e') or {}).get('value'), d['config']['parallelism'][:60])
except Exception as e: print('no line',e)
"
  tail -n 2 $O/bench_$tag.err | cut -c1-300
}
EXTRA="" run default_with_e2e DET_PEER_MINB=1
EXTRA="--no-e2e" run minb4 DET_PEER_MINB=4
EXTRA="--no-e2e" run minb1 DET_PEER_MINB=1
