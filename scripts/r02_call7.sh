#!/bin/bash
# Round 2, GPU call 7 (gpurun --gpus 8): the headline step on 8 GPUs with the owner-side exchange and with the round-1
# remote-probe kernels (in-run parity at N=8 in both), BASELINE configs[3] (--workload c4: 1B keys, dim 128) and
# configs[4] (--workload c5: sharded forward + backward).
set -u
export DET_NO_REBUILD=1
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c7_n$N
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
nvidia-smi topo -m > $O/topo.txt 2>&1
run() {  # tag, extra args...
  tag=$1; shift
  DET_XCHG_TIMING=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 \
    bench.py --gpus $N --warmup 10 --no-e2e "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "bench N=$N $tag exit $?"
  grep "det xchg timing rank 0" $O/bench_$tag.err
  tail -n 1 $O/bench_$tag.json | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read())
  print({k:d.get(k) for k in ('metric','value','unit','ms_per_step','find_ms','insert_ms')}, (d.get('parity') or {}).get('mismatches'), (d.get('parity') or {}).get('checked'), (d.get('roofline_nvlink') or {}).get('frac'), (d.get('no_exchange') or {}).get('value'))
except Exception as e: print('no line',e)
"
  tail -n 2 $O/bench_$tag.err | cut -c1-300
  lap $tag
}
run push --steps 200 --exchange push
run peer --steps 200 --exchange peer
run c4_push --steps 100 --exchange push --workload c4
run c5 --steps 30 --workload c5
