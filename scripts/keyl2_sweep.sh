#!/bin/bash
# Round-2 experiment: L2 prefetch-size qualifier on the 64 B bucket loads of every probe (-DDET_KEY_L2=64|128:
# SASS LDG.E.LTC64B / LTC128B instead of the unqualified default).  ncu (profiles/r01_find_kernel_tma_dim64.csv) shows
# ~125 B of DRAM read per probe; if the L2 fills 128 B lines for the unqualified loads, LTC64B halves that (= -10 % of
# find's traffic).  Per variant: microbench of find / insert_existing at dim 64 and the DRAM bytes of one find launch.
# The default library is restored at the end.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/keyl2_sweep.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # $1 = tag, $2 = extra nvcc flags
  DET_NVCC_EXTRA="$2" python -m recommenders_addons_b200.build --force > /dev/null 2>&1
  timeout 600 python scripts/microbench.py --ops find,insert_existing --dims 64 --resident 50000000 --tag "$1" >> gpurun_out/keyl2_sweep.jsonl 2>> gpurun_out/keyl2_sweep.err
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:find_kernel_tma" -c 3 --csv \
    --log-file "gpurun_out/keyl2_$1.csv" python scripts/microbench.py --ops find --dims 64 --resident 50000000 --reps 1 --tag "$1" > /dev/null 2>> gpurun_out/keyl2_sweep.err
  tail -n 1 "gpurun_out/keyl2_$1.csv" | sed "s/^/$1: /"
}
: > gpurun_out/keyl2_sweep.jsonl
run default ""
run ltc64 "-DDET_KEY_L2=64"
run ltc128 "-DDET_KEY_L2=128"
python -m recommenders_addons_b200.build --force > /dev/null 2>&1
cut -c1-260 gpurun_out/keyl2_sweep.jsonl
