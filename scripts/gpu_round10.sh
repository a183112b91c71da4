#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_table_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 500 -k "host" > gpurun_out/pytest_host.log 2>&1
echo "pytest host exit $?"; tail -n 3 gpurun_out/pytest_host.log
timeout 600 python bench.py --no-cpu-baseline --steps 200 --warmup 10 > gpurun_out/bench10.json 2> gpurun_out/bench10.err
echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench10.json').read()); print(d['value'], json.dumps(d['e2e']))"; tail -n 3 gpurun_out/bench10.err
