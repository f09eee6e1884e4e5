#!/bin/bash
# round-2 third GPU pass: suite, stamped probes after the interleaved split / NM template, alt (768-thread) A/B, ncu captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2c_pytest.log
P=gpurun_out/r2c_probe.jsonl; E=gpurun_out/r2c_probe.err; : > $P; : > $E
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
VLCAL_PK_KPT=4 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag kpt4 >> $P 2>> $E
VLCAL_PK_ATOM=1 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --tag atom1 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 >> $P 2>> $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag t768 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag t768 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
cat $P; tail -n 5 $E
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2c_pk_c2 python tools/pk_probe.py --config C2 --modes 3 --reps 1 > gpurun_out/r2c_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; grep '"what": "solve"' gpurun_out/r2c_ncu_c2.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2c_pk_c3 python tools/pk_probe.py --config C3 --modes 3 --reps 1 > gpurun_out/r2c_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"; grep '"what": "solve"' gpurun_out/r2c_ncu_c3.log | cut -c1-400
