#!/bin/bash
# Round 2, last single-GPU check after the two small c3 fusions (unique flag + block sums in one pass; the one-id-per-row
# check inside lookup_identity_kernel): GPU suite, smoke(), the c3 line.
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02last
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee $O/tests_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee $O/smoke.log
timeout 400 python bench.py --workload c3 --steps 100 --warmup 5 > $O/c3.json 2> $O/c3.err
python - <<'P'
import json
try:
  d=json.loads(open('gpurun_out/r02last/c3.json').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('ms_per_step','phases_ms','host_syncs_per_step','parity')}, d['roofline']['frac'])
except Exception as e: print('no line', e)
P
tail -n 2 $O/c3.err | cut -c1-200
