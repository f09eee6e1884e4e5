#!/bin/bash
# round-2 sixth GPU pass (1 GPU): packed-fp32 classifier (FFMA2) as default, scalar build as alt; suite, probes, ncu, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
P=gpurun_out/r2g_probe.jsonl; E=gpurun_out/r2g_probe.err; : > $P; : > $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag scalar_fp32 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag scalar_fp32 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 --tag scalar_fp32 >> $P 2>> $E
cat $P | cut -c1-900; tail -n 5 $E
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench_c3.json 2> gpurun_out/r2g_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-400 gpurun_out/r2g_bench_c3.json
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err
echo "bench c2 rc=$?"; cut -c1-400 gpurun_out/r2g_bench_c2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2g_pk_c2 python tools/pk_probe.py --config C2 --modes 3 --reps 1 > gpurun_out/r2g_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2g_pk_c3 python tools/pk_probe.py --config C3 --modes 3 --reps 1 > gpurun_out/r2g_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
