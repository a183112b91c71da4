#!/bin/bash
# First GPU action of round 2: run the capacity-management (csrc/evict.cu) and restrict-policy tests, which were
# written after round 1's GPU budget was spent and are skipped without DET_TEST_UNVALIDATED=1.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_validate_evict.sh'
# Outputs land in gpurun_out/evict_*.log.  When they are green: drop the skipif gates in tests/test_evict_gpu.py,
# tests/test_restrict_gpu.py and tests/test_spill_gpu.py, and record the run under profiles/.
set -u
mkdir -p gpurun_out
export DET_TEST_UNVALIDATED=1
timeout 900 python -m pytest tests/test_evict_gpu.py tests/test_restrict_gpu.py tests/test_spill_gpu.py tests/test_callers_gpu.py tests/test_segreduce_gpu.py -q -m gpu 2>&1 | tee gpurun_out/evict_tests.log | tail -40
# memory checker over the small cases (every kernel of evict.cu runs at least once)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_evict_gpu.py -q -m gpu \
  -k "basic or lfu or custom or growth or touch" > gpurun_out/evict_memcheck.log 2>&1
echo "memcheck exit: $?" | tee -a gpurun_out/evict_tests.log
tail -5 gpurun_out/evict_memcheck.log
# racecheck on shared-memory histograms / counters
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_evict_gpu.py -q -m gpu \
  -k "explicit" > gpurun_out/evict_racecheck.log 2>&1
echo "racecheck exit: $?" | tee -a gpurun_out/evict_tests.log
# the validated suite must be untouched by the new translation unit
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/evict_tests.log
# the A/B experiments and microbenches that used to follow live in scripts/round2_sweeps.sh (a separate gpurun call)
# the opt-in paths of the Python mirror on the VALIDATED suites: deterministic gradient dedupe (det_segment_reduce) and
# the trainable sparse lookup through det_sparse_segment_sum; green here -> make both the defaults (scripts/README.md)
DET_GRAD_REDUCE=det DET_SPARSE_TRAIN_FUSED=1 timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_peer_gpu.py tests/test_table_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/evict_tests.log
