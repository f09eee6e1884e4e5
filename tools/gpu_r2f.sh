#!/bin/bash
# round-2 fifth GPU pass (1 GPU): suite + probes + both bench arms with the prefix-sum deferral build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
P=gpurun_out/r2f_probe.jsonl; E=gpurun_out/r2f_probe.err; : > $P; : > $E
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
cat $P | cut -c1-700; tail -n 5 $E
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-400 gpurun_out/r2f_bench_c3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_c3_ref.json 2> gpurun_out/r2f_bench_c3_ref.err
echo "ref c3 rc=$?"; cut -c1-300 gpurun_out/r2f_bench_c3_ref.json
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
echo "bench c2 rc=$?"; cut -c1-400 gpurun_out/r2f_bench_c2.json
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
echo "bench c5 rc=$?"; cut -c1-400 gpurun_out/r2f_bench_c5.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2f_smoke.log
