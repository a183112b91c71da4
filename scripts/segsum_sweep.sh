#!/bin/bash
# Round-2 experiment on the fused sparse-lookup gather (segment_sum_kernel<4>: 4.2 TB/s in round 1, 4 segments per
# lane-group in lock step = 4 row loads in flight per lane, registers capped for 3 CTAs/SM): rebuild libdetable.so with
# more segments per lane-group / a different register cap and time det_lookup_sparse with 1 id and 4 ids per row.
# The default library is restored at the end.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/segsum_sweep.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # $1 = tag, $2 = extra nvcc flags
  DET_NVCC_EXTRA="$2" python -m recommenders_addons_b200.build --force -v 2>&1 | grep -A2 "segment_sum_kernelILi4ELb0" | grep "Used" | sed "s/^/$1: /" | tee -a gpurun_out/segsum_sweep.log
  timeout 600 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 64 --resident 50000000 --tag "$1" >> gpurun_out/segsum_sweep.jsonl 2>> gpurun_out/segsum_sweep.err
}
: > gpurun_out/segsum_sweep.jsonl
run default ""
run u3_minb4 "-DDET_SEG_U=3 -DDET_SEG_MINB=4"
run u6_minb2 "-DDET_SEG_U=6 -DDET_SEG_MINB=2"
python -m recommenders_addons_b200.build --force > /dev/null 2>&1
# the cp.async-staged variant (segment_sum_staged_kernel, default library): correctness on the GPU suite first, then timing
DET_SEGSUM_STAGED=1 timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "sparse" 2>&1 | tail -2 | tee -a gpurun_out/segsum_sweep.log
DET_SEGSUM_STAGED=1 timeout 600 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 16,64,128 --resident 50000000 --tag staged >> gpurun_out/segsum_sweep.jsonl 2>> gpurun_out/segsum_sweep.err
timeout 600 python scripts/microbench.py --ops lookup_sparse_1id,lookup_sparse_4ids --dims 16,64,128 --resident 50000000 --tag plain >> gpurun_out/segsum_sweep.jsonl 2>> gpurun_out/segsum_sweep.err
cut -c1-260 gpurun_out/segsum_sweep.jsonl
