#!/bin/bash
# round-2 second GPU pass: GPU suite with the new tests, stamped probes (per-block times), bench smoke (both arms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf --durations=8 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/r2b_pytest.log
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps > gpurun_out/r2b_probe.jsonl 2> gpurun_out/r2b_probe.err
VLCAL_PK_KPT=2 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag kpt2 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag t768 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
VLCAL_LIB=$ALT VLCAL_PK_KPT=2 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag t768_kpt2 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag t768 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
VLCAL_PK_TMA=1 VLCAL_PK_KPT=2 timeout 300 python tools/pk_probe.py --config C2 --modes 1,3 --stamps --tag tma_kpt2 >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
VLCAL_PK_DYNAMIC=0 timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag static >> gpurun_out/r2b_probe.jsonl 2>> gpurun_out/r2b_probe.err
cat gpurun_out/r2b_probe.jsonl
tail -5 gpurun_out/r2b_probe.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
echo "bench rc=$?"; cat gpurun_out/r2b_bench_c3.json; tail -5 gpurun_out/r2b_bench_c3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_c3_ref.json 2> gpurun_out/r2b_bench_c3_ref.err
echo "bench ref rc=$?"; cat gpurun_out/r2b_bench_c3_ref.json; tail -5 gpurun_out/r2b_bench_c3_ref.err
timeout 600 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
echo "bench c2 rc=$?"; cat gpurun_out/r2b_bench_c2.json; tail -5 gpurun_out/r2b_bench_c2.err
