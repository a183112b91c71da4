#!/bin/bash
# Round 2, GPU call 3 (gpurun --gpus N, N = 2 by default): the owner-side exchange on real NVLink -- the multi-process
# parity test, then bench.py at N with --exchange push (default) / peer (round-1 remote-probe kernels) / nccl.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r02_call3.sh 2'
set -u
export DET_NO_REBUILD=1
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c3_n$N
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 12 | tee $O/test_multigpu.log
  lap multigpu-test
fi
for EX in ${EXCHANGES:-push peer nccl}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps ${STEPS:-300} --warmup 10 --exchange $EX --no-e2e > $O/bench_$EX.json 2> $O/bench_$EX.err
  echo "bench N=$N $EX exit $?"
  tail -n 1 $O/bench_$EX.json | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read())
  print({k:d.get(k) for k in ('value','ms_per_step','find_ms','insert_ms','parity','roofline_nvlink','no_exchange')})
except Exception as e: print('no line',e)
"
  tail -n 4 $O/bench_$EX.err | cut -c1-300
  lap bench-$EX
done
