#!/bin/bash
# Round 2, final single-GPU validation: the GPU suite, smoke(), the default bench line (parity, hard cases, c3, e2e, CPU
# arm), DRAM traffic of the headline kernel with the FINAL sources (profiles/r02_traffic.json), the ncu launch list of the
# same bench command, one `ncu --set full` capture of the top kernels.
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02final
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee $O/tests_all.log
lap tests
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/smoke.log
lap smoke
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench exit $?"
python - <<'P'
import json
try:
  d=json.loads(open('gpurun_out/r02final/bench_n1.json').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('value','ms_per_step','find_Mkeys_s','insert_Mkeys_s','parity')})
  print('e2e', {k:d['e2e'].get(k) for k in ('value','drained_every_step_value','sequential_value','pcie_frac','prefetch_rows_checked','pipelined_error')})
  print('c3', {k:d['c3'].get(k) for k in ('ms_per_step','phases_ms','host_syncs_per_step','parity','error')}, d['c3'].get('roofline',{}).get('frac'))
  print('hard', {k:round(v['Mkeys_s']) for k,v in d['hard_cases'].items() if isinstance(v,dict) and 'Mkeys_s' in v})
  print('cpu', {k:d['cpu_baseline'].get(k) for k in ('value','cores','value_min','value_max')}, 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
except Exception as e: print('no line', e)
P
tail -n 4 $O/bench_n1.err | cut -c1-300
lap bench
timeout 400 python scripts/ncu_traffic.py 2>&1 | tail -n 2 | cut -c1-400
lap traffic
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:det::" -c 2500 --csv --log-file $O/launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-hard-cases > $O/bench_under_ncu.log 2>&1
echo "ncu launch list exit $?"; python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02final/launches_bench_n1.csv')))
hdr=None; seq=[]
for r in rows:
    if r and r[0]=='ID': hdr=r
    elif hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); seq.append((d['Kernel Name'][:48], float(d['Metric Value'].replace(',',''))))
print(len(seq), 'launches')
# the headline timed steps: the last find/insert pairs before the c3 table is filled
idx=[i for i,(n,_) in enumerate(seq) if 'find_kernel_tma' in n]
if idx:
    j=idx[-1]
    for n,t in seq[max(0,j-4):j+3]: print("%-48s %8.1f us"%(n,t/1e3))
P
lap launch-list
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:find_kernel_tma|insert_kernel_tma" --launch-skip 100 -c 2 -f -o $O/top_kernels_c2 \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-hard-cases --no-c3 > $O/bench_under_ncu_full.log 2>&1
echo "ncu --set full (find / insert) exit $?"
python scripts/ncu_summary.py $O/top_kernels_c2.ncu-rep > $O/top_kernels_c2_summary.csv 2> $O/top_kernels_summary.err; cut -c1-300 $O/top_kernels_c2_summary.csv | head -6
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:lookup_identity_kernel|segment_reduce_kernel|apply_staged_kernel" --launch-skip 9 -c 3 -f -o $O/top_kernels_c3 \
  python bench.py --workload c3 --steps 2 --warmup 2 > $O/bench_c3_under_ncu_full.log 2>&1
echo "ncu --set full (c3 kernels) exit $?"
python scripts/ncu_summary.py $O/top_kernels_c3.ncu-rep > $O/top_kernels_c3_summary.csv 2>> $O/top_kernels_summary.err; cut -c1-300 $O/top_kernels_c3_summary.csv | head -6
lap ncu-full
