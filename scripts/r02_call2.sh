#!/bin/bash
# Round 2, GPU call 2 (1 GPU): the new bench line end to end (parity, hard cases, c3, reference arm), the variant sweep
# that call 1 lost to an ABI bump, the kernel-by-kernel launch list of the c3 step, traffic of the headline kernel.
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c2
mkdir -p $O
V=recommenders_addons_b200/lib/variants
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }
timeout 900 python bench.py --steps 300 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench exit $?"; cut -c1-3000 $O/bench_n1.json; tail -n 5 $O/bench_n1.err
lap bench
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
echo "reference exit $?"; cut -c1-1500 $O/bench_ref.json; tail -n 3 $O/bench_ref.err
lap reference
for tag in default ltc64 ltc128 minb5 minb6 minb8; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 python scripts/microbench.py --ops find,insert_existing --dims 64 --resident 50000000 --tag $tag 2>> $O/sweep.err | grep -v find_exists >> $O/sweep.jsonl
done
cut -c1-200 $O/sweep.jsonl
for tag in default ltc64; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:find_kernel_tma" -c 12 --csv \
    --log-file $O/keyl2_$tag.csv python scripts/microbench.py --ops find --dims 64 --resident 50000000 --reps 1 --tag $tag > /dev/null 2>> $O/sweep.err
done
lap sweep
for tag in default segu3 segu6; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 64 --resident 50000000 --tag $tag >> $O/segsum.jsonl 2>> $O/segsum.err
done
cut -c1-200 $O/segsum.jsonl
lap segsum
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/c3_launches.csv \
  python bench.py --workload c3 --steps 2 --warmup 2 > $O/c3_under_ncu.json 2> $O/c3_under_ncu.err
tail -n 60 $O/c3_launches.csv | cut -d, -f5,12- | cut -c1-160
lap c3-launches
timeout 600 python scripts/ncu_traffic.py 2>&1 | tail -n 3
lap traffic
timeout 300 python scripts/evict_microbench.py --capacity 100000000 --steps 100 > $O/evict_microbench.jsonl 2> $O/evict_microbench.err
tail -n 6 $O/evict_microbench.jsonl | cut -c1-300
timeout 300 python scripts/spill_microbench.py > $O/spill_microbench.jsonl 2> $O/spill_microbench.err
cut -c1-300 $O/spill_microbench.jsonl
lap evict+spill
tail -n 3 $O/*.err | cut -c1-300
