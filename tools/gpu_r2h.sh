#!/bin/bash
# round-2 seventh GPU pass (1 GPU): every block finalizes (warp NID, triple-buffered accumulators), zeroing overlapped with the
# simplex step; alt = 512 threads per block, 4 points per lane (two packed pairs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
P=gpurun_out/r2h_probe.jsonl; E=gpurun_out/r2h_probe.err; : > $P; : > $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
VLCAL_LIB=$ALT VLCAL_PK_KPT=4 timeout 300 python tools/pk_probe.py --config C2 --modes 1,3 --stamps --tag t512_k4 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag t512_k2 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> $P 2>> $E
VLCAL_LIB=$ALT VLCAL_PK_KPT=4 timeout 600 python tools/pk_probe.py --config C3 --modes 1,3 --stamps --reps 3 --tag t512_k4 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag t512_k2 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
VLCAL_LIB=$ALT VLCAL_PK_KPT=4 timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 --tag t512_k4 >> $P 2>> $E
cat $P | cut -c1-900; tail -n 5 $E
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-300 gpurun_out/r2h_bench_c3.json
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err
echo "bench c2 rc=$?"; cut -c1-300 gpurun_out/r2h_bench_c2.json
