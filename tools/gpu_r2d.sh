#!/bin/bash
# round-2 fourth GPU pass: 768-thread default + per-tile deferral + rank sort; alt = 3 x 256; bench both configs; ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest.log
P=gpurun_out/r2d_probe.jsonl; E=gpurun_out/r2d_probe.err; : > $P; : > $E
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> $P 2>> $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag b256x3 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag b256x3 >> $P 2>> $E
VLCAL_PK_TMA=1 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag tma >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
cat $P; tail -n 5 $E
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-600 gpurun_out/r2d_bench_c3.json
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err
echo "bench c2 rc=$?"; cut -c1-600 gpurun_out/r2d_bench_c2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2d_pk_c2 python tools/pk_probe.py --config C2 --modes 3 --reps 1 > gpurun_out/r2d_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2d_pk_c3 python tools/pk_probe.py --config C3 --modes 3 --reps 1 > gpurun_out/r2d_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
