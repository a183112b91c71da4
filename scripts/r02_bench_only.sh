#!/bin/bash
# the default bench line of HEAD (what the driver runs at N=1)
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/r02final
timeout 600 python bench.py > gpurun_out/r02final/bench_n1_head.json 2> gpurun_out/r02final/bench_n1_head.err
echo "bench exit $?"
python - <<'P'
import json
try:
  d=json.loads(open('gpurun_out/r02final/bench_n1_head.json').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('value','ms_per_step','find_Mkeys_s','insert_Mkeys_s')}, d['parity']['mismatches'])
  print('e2e', d['e2e']['value'], d['e2e'].get('pcie_frac'), 'c3', d['c3'].get('ms_per_step'), d['c3'].get('host_syncs_per_step'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['value_min'], d['cpu_baseline']['value_max'], 'traffic', d['roofline']['traffic'])
except Exception as e: print('no line', e)
P
tail -n 3 gpurun_out/r02final/bench_n1_head.err | cut -c1-200
