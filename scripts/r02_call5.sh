#!/bin/bash
# Round 2, GPU call 5 (1 GPU): the GPU suite after the round's kernel changes, the c3 step with the one-call backward and
# the re-scheduled long-group path (bench line + kernel launch list), what one eviction event is made of.
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c5
mkdir -p $O
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 | tee $O/tests_all.log
lap tests
timeout 400 python bench.py --workload c3 --steps 50 --warmup 5 > $O/c3_det.json 2> $O/c3_det.err
cut -c1-900 $O/c3_det.json; tail -n 3 $O/c3_det.err
lap c3
timeout 300 python scripts/microbench.py --ops segment_reduce,index_add --dims 64 --resident 20000000 > $O/segment_reduce.jsonl 2> $O/segment_reduce.err
cut -c1-230 $O/segment_reduce.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/c3_launches.csv \
  python bench.py --workload c3 --steps 2 --warmup 2 > $O/c3_under_ncu.json 2> $O/c3_under_ncu.err
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02c5/c3_launches.csv')))
hdr=None; by={}; order=[]
for r in rows:
    if r and r[0]=='ID': hdr=r
    elif hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['ID'] not in by: order.append(d['ID']); by[d['ID']]={'name':d['Kernel Name'][:70],'grid':d['Grid Size']}
        by[d['ID']][d['Metric Name']]=float(d['Metric Value'].replace(',',''))
idxs=[i for i,k in enumerate(order) if 'resolve_slots' in by[k]['name']]
if len(idxs)>=2:
    tot=0
    for k in order[idxs[-2]:idxs[-1]]:
        b=by[k]; tot+=b['gpu__time_duration.sum']
        print("%-70s %-10s %8.1f us rd %7.1f wr %7.1f MB"%(b['name'],b['grid'],b['gpu__time_duration.sum']/1e3,b.get('dram__bytes_read.sum',0)/1e6,b.get('dram__bytes_write.sum',0)/1e6))
    print('total us',tot/1e3)
P
lap c3-launches
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/evict_launches.csv \
  python scripts/evict_microbench.py --capacity 100000000 --only-events > $O/evict_events.jsonl 2> $O/evict_events.err
python - <<'P'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02c5/evict_launches.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r and r[0]=='ID': hdr=r
    elif hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); n=d['Kernel Name'][:60]
        if any(x in n for x in ('repair','minmax','hist','pick','evict','select','purge','classify','read')):
            a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(d['Metric Value'].replace(',',''))
for n,(c,t) in agg.items(): print("%-60s x%-4d %9.1f us total"%(n,c,t/1e3))
P
cat $O/evict_events.jsonl | cut -c1-200
lap evict-launches
timeout 300 python scripts/evict_microbench.py --capacity 100000000 --steps 200 > $O/evict_microbench.jsonl 2> $O/evict_microbench.err
tail -n 2 $O/evict_microbench.jsonl | cut -c1-300
lap evict
