#!/bin/bash
# ncu launch list of the default bench command (short run): every kernel of this repo with its duration, cold-cache and
# serialised -- only each kernel's SHARE of a step is comparable with the live CUDA-event numbers of bench.py.
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02final
mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k "regex:find_kernel|insert_kernel|lookup_identity|resolve_slots|segment_sum|segment_offsets|unique_|radix_|group_|segment_reduce|apply_staged|scan_" \
  -c 3000 --csv --log-file $O/launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-hard-cases > $O/bench_under_ncu.log 2>&1
echo "ncu launch list exit $?"
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02final/launches_bench_n1.csv')))
hdr=None; seq=[]
for r in rows:
    if r and r[0]=='ID': hdr=r
    elif hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); seq.append((d['Kernel Name'][:60], float(d['Metric Value'].replace(',',''))))
print(len(seq), 'launches')
idx=[i for i,(n,_) in enumerate(seq) if 'find_kernel_tma' in n]
if idx:
    j=idx[-2] if len(idx)>1 else idx[-1]
    for n,t in seq[max(0,j-3):j+3]: print("%-60s %8.1f us"%(n,t/1e3))
print('... c3 step:')
idx=[i for i,(n,_) in enumerate(seq) if 'lookup_identity' in n]
if len(idx)>=2:
    for n,t in seq[idx[-2]:idx[-1]]: print("%-60s %8.1f us"%(n,t/1e3))
P
