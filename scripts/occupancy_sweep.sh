#!/bin/bash
# Round-2 experiment on the headline kernels (find_kernel_tma / insert_kernel_tma, 64 registers, 4 CTAs/SM today):
# rebuild libdetable.so with a minimum of n resident CTAs per SM (register cap) and time find / insert with resident
# keys at dim 64.  The default library is restored at the end.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/occupancy_sweep.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # $1 = tag, $2 = extra nvcc flags
  DET_NVCC_EXTRA="$2" python -m recommenders_addons_b200.build --force -v 2>&1 | grep -A2 "find_kernel_tmaILi16\|insert_kernel_tmaILi16ELb0" | grep "Used" | sed "s/^/$1: /" | tee -a gpurun_out/occupancy_sweep.log
  timeout 600 python scripts/microbench.py --ops find,insert_existing --dims 64 --resident 50000000 --tag "$1" >> gpurun_out/occupancy_sweep.jsonl 2>> gpurun_out/occupancy_sweep.err
}
: > gpurun_out/occupancy_sweep.jsonl
run default ""
run minb5 "-DDET_FIND_MINB=5 -DDET_INSERT_MINB=5"
run minb6 "-DDET_FIND_MINB=6 -DDET_INSERT_MINB=6"
run minb8 "-DDET_FIND_MINB=8 -DDET_INSERT_MINB=8"
python -m recommenders_addons_b200.build --force > /dev/null 2>&1
cut -c1-260 gpurun_out/occupancy_sweep.jsonl
