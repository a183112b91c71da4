#!/bin/bash
# Round 2, SECOND GPU call (after scripts/gpu_validate_evict.sh is green): every measurement that decides a follow-up,
# each under its own timeout, results in gpurun_out/ (copy what is kept to profiles/r02_*).  ~25 min when nothing hangs.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash scripts/round2_sweeps.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export DET_TEST_UNVALIDATED=1
# cost of the unvalidated features once they are green (short runs)
timeout 400 python scripts/evict_microbench.py --capacity 20000000 --steps 100 > gpurun_out/evict_microbench.jsonl 2> gpurun_out/evict_microbench.err
echo "evict microbench exit: $?"; tail -n 3 gpurun_out/evict_microbench.jsonl
timeout 400 python scripts/spill_microbench.py > gpurun_out/spill_microbench.jsonl 2> gpurun_out/spill_microbench.err
echo "spill microbench exit: $?"; cat gpurun_out/spill_microbench.jsonl
# round-2 candidate: batched slot claims for inserts of NEW keys (DET_CLAIM_BATCH=1, common.cuh); correctness first
# (the validated table suite under the variant), then the A/B of insert_new / insert_existing at dim 16 / 64 / 128
DET_CLAIM_BATCH=1 timeout 900 python -m pytest tests/test_table_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/claim_batch_tests.log
timeout 600 python scripts/microbench.py --ops insert_new,insert_existing --resident 20000000 --tag serial > gpurun_out/claim_serial.jsonl 2> gpurun_out/claim_serial.err
DET_CLAIM_BATCH=1 timeout 600 python scripts/microbench.py --ops insert_new,insert_existing --resident 20000000 --tag batched > gpurun_out/claim_batched.jsonl 2> gpurun_out/claim_batched.err
echo "claim A/B:"; cat gpurun_out/claim_serial.jsonl gpurun_out/claim_batched.jsonl | cut -c1-300
# how much DRAM traffic does one random 32 / 64 / 128 / 256 B read cost? (decides the bucket-width follow-up)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__sectors_read.sum --clock-control none -k "regex:[iI]ndex|gather" --csv \
  --log-file gpurun_out/granularity.csv python scripts/probe_granularity.py > gpurun_out/granularity.json 2> gpurun_out/granularity.err
echo "granularity probe exit: $?"; grep -ci "index\|gather" gpurun_out/granularity.csv
L2_FETCH=32 timeout 600 ncu --metrics dram__bytes_read.sum,dram__sectors_read.sum --clock-control none -k "regex:[iI]ndex|gather" --csv \
  --log-file gpurun_out/granularity_l2fetch32.csv python scripts/probe_granularity.py > gpurun_out/granularity_l2fetch32.json 2>> gpurun_out/granularity.err
# round-2 candidate: deterministic per-unique gradient sum (det_segment_reduce) vs torch index_add in the c3 step
timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --grad-reduce det > gpurun_out/c3_det.json 2> gpurun_out/c3_det.err
timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --grad-reduce torch > gpurun_out/c3_torch.json 2> gpurun_out/c3_torch.err
timeout 400 python scripts/microbench.py --ops segment_reduce,index_add --dims 16,64,128 --resident 20000000 > gpurun_out/segment_reduce.jsonl 2> gpurun_out/segment_reduce.err
for c in 37 74 296; do
  DET_SEGRED_LONG_CTAS=$c timeout 300 python scripts/microbench.py --ops segment_reduce --dims 64 --resident 20000000 --tag "long_ctas_$c" >> gpurun_out/segment_reduce.jsonl 2>> gpurun_out/segment_reduce.err
done
cut -c1-260 gpurun_out/segment_reduce.jsonl
echo "c3 A/B (det_segment_reduce vs index_add):"; cut -c1-260 gpurun_out/c3_det.json gpurun_out/c3_torch.json
# L2 prefetch-size qualifier on the bucket loads (the ~125 B of DRAM read per probe), register caps, segment_sum variants
timeout 900 bash scripts/keyl2_sweep.sh 2>&1 | tail -8
timeout 900 bash scripts/occupancy_sweep.sh 2>&1 | tail -6
timeout 900 bash scripts/segsum_sweep.sh 2>&1 | tail -6
