#!/bin/bash
# Round 2, GPU call 1 (1 GPU, ~18 min): the whole ungated GPU suite, the opt-in paths on the validated suites, then every
# A/B measurement that decides a follow-up.  Variant libraries are PREBUILT in the authoring container
# (python -m recommenders_addons_b200.build --variant <tag> <flags>; lib/variants/*.so travel with the snapshot) and selected
# with DET_LIB_PATH, so the box spends no GPU-minutes compiling.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r02_call1.sh'
set -u
export DET_NO_REBUILD=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c1
mkdir -p $O
V=recommenders_addons_b200/lib/variants
t0=$(date +%s)
lap() { echo "== $1 done at +$(( $(date +%s) - t0 )) s"; }

timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee $O/tests_all.log
lap tests
DET_GRAD_REDUCE=det DET_SPARSE_TRAIN_FUSED=1 timeout 400 python -m pytest tests/test_fused_gpu.py tests/test_peer_gpu.py tests/test_callers_gpu.py tests/test_table_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee $O/tests_optin.log
lap optin
DET_CLAIM_BATCH=1 timeout 400 python -m pytest tests/test_table_gpu.py tests/test_fused_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $O/tests_claim_batch.log
DET_SEGSUM_STAGED=1 timeout 300 python -m pytest tests/test_fused_gpu.py -q -m gpu -k sparse -p no:cacheprovider 2>&1 | tail -3 | tee $O/tests_segsum_staged.log
lap variant-tests

# batched slot claims: new-key inserts (the dynamic part of a dynamic embedding)
timeout 300 python scripts/microbench.py --ops insert_new,insert_existing --dims 16,64,128 --resident 20000000 --tag serial > $O/claim.jsonl 2> $O/claim.err
DET_CLAIM_BATCH=1 timeout 300 python scripts/microbench.py --ops insert_new,insert_existing --dims 16,64,128 --resident 20000000 --tag batched >> $O/claim.jsonl 2>> $O/claim.err
cut -c1-230 $O/claim.jsonl
lap claim

# DRAM fetch granularity of random reads
timeout 300 ncu --metrics dram__bytes_read.sum,dram__sectors_read.sum --clock-control none -k "regex:[iI]ndex|gather" --csv \
  --log-file $O/granularity.csv python scripts/probe_granularity.py > $O/granularity.json 2> $O/granularity.err
grep -c "dram__bytes_read" $O/granularity.csv
lap granularity

# L2 prefetch qualifier on the bucket loads + register caps: find / insert at dim 64, DRAM bytes of one find launch
for tag in default ltc64 ltc128 minb5 minb6 minb8; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 python scripts/microbench.py --ops find,insert_existing --dims 64 --resident 50000000 --tag $tag 2>> $O/sweep.err | grep -v find_exists >> $O/sweep.jsonl
done
for tag in default ltc64 ltc128; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:find_kernel_tma" -c 3 --csv \
    --log-file $O/keyl2_$tag.csv python scripts/microbench.py --ops find --dims 64 --resident 50000000 --reps 1 --tag $tag > /dev/null 2>> $O/sweep.err
  tail -n 3 $O/keyl2_$tag.csv | cut -c1-200 | sed "s/^/$tag: /"
done
cut -c1-230 $O/sweep.jsonl
lap keyl2+occupancy

# fused sparse lookup gather: segments per lane-group / staged variant
for tag in default segu3 segu6; do
  lib=""; [ $tag != default ] && lib="$PWD/$V/libdetable_$tag.so"
  DET_LIB_PATH=$lib timeout 200 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 64 --resident 50000000 --tag $tag >> $O/segsum.jsonl 2>> $O/segsum.err
done
DET_SEGSUM_STAGED=1 timeout 200 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 16,64,128 --resident 50000000 --tag staged >> $O/segsum.jsonl 2>> $O/segsum.err
cut -c1-230 $O/segsum.jsonl
lap segsum

# gradient dedupe and the c3 step
timeout 300 python scripts/microbench.py --ops segment_reduce,index_add --dims 16,64,128 --resident 20000000 > $O/segment_reduce.jsonl 2> $O/segment_reduce.err
cut -c1-230 $O/segment_reduce.jsonl
timeout 400 python bench.py --workload c3 --steps 30 --warmup 5 --grad-reduce det > $O/c3_det.json 2> $O/c3_det.err
timeout 400 python bench.py --workload c3 --steps 30 --warmup 5 --grad-reduce torch > $O/c3_torch.json 2> $O/c3_torch.err
cut -c1-400 $O/c3_det.json $O/c3_torch.json
lap c3

# capacity management and spill: what they cost
timeout 300 python scripts/evict_microbench.py --capacity 100000000 --steps 100 > $O/evict_microbench.jsonl 2> $O/evict_microbench.err
tail -n 6 $O/evict_microbench.jsonl | cut -c1-300
timeout 300 python scripts/spill_microbench.py > $O/spill_microbench.jsonl 2> $O/spill_microbench.err
cut -c1-300 $O/spill_microbench.jsonl
lap evict+spill
tail -n 3 $O/*.err | cut -c1-300
