#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench (both kernel variants) + ncu launch list + one full ncu capture.
# usage: tools/gpu_check.sh <tag> [micro]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/${TAG}_pytest_gpu.log
if [ "$2" == "micro" ]; then echo "== microbench"; timeout 300 tools/microbench > $OUT/${TAG}_microbench.log 2>&1; cat $OUT/${TAG}_microbench.log; fi
echo "== bench (default kernel)"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench (host-driven loop)"; timeout 900 python bench.py --steps 10 --warmup 3 --solver host --no-cpu-baseline > $OUT/${TAG}_bench_hostloop.json 2> $OUT/${TAG}_bench_hostloop.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench_hostloop.json
echo "== bench (exact fp64 kernel)"; timeout 900 python bench.py --steps 5 --warmup 3 --variant 1 --no-cpu-baseline > $OUT/${TAG}_bench_exact.json 2> $OUT/${TAG}_bench_exact.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench_exact.json
echo "== bench (filter, 2 points/thread)"; timeout 900 python bench.py --steps 5 --warmup 3 --variant 2 --no-cpu-baseline > $OUT/${TAG}_bench_kpt2.json 2> $OUT/${TAG}_bench_kpt2.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench_kpt2.json
echo "== kernel scaling"; timeout 600 python tools/kernel_scaling.py > $OUT/${TAG}_kernel_scaling.jsonl 2>&1; cat $OUT/${TAG}_kernel_scaling.jsonl
echo "== configs (skipped unless CONFIGS=1)"; [ "$CONFIGS" == "1" ] && timeout 1500 python tests/perf/config_bench.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; echo "rc=$?"; cut -c1-420 $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
echo "== K3 value+gradient kernel"; timeout 300 python tests/perf/grad_bench.py > $OUT/${TAG}_grad_bench.json 2> $OUT/${TAG}_grad_bench.err; tail -c 600 $OUT/${TAG}_grad_bench.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>&1; tail -c 400 $OUT/${TAG}_bench_ref.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 600 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:nid_hist -s 30 -c 2 -f -o $OUT/${TAG}_nid_hist python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la $OUT | tail -20
