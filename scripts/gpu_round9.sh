#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"apply_staged|accum_kernel|unique_|partition_|remove_kernel|export_write|export_count|peer_find|peer_insert|peer_route|insert_kernel_tma" \
  -s 8 -c 40 -o gpurun_out/prof_misc python scripts/ncu_misc.py > gpurun_out/ncu_misc.log 2>&1
echo "ncu misc exit $?"; tail -n 3 gpurun_out/ncu_misc.log
