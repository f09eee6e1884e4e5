#!/bin/bash
# round-2 last GPU pass (1 GPU): the suite with the new bin-count / finalizer-round test, one default bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout 600 -rf > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2o_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2o_bench_c3.json 2> gpurun_out/r2o_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-200 gpurun_out/r2o_bench_c3.json
