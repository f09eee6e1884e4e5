#!/bin/bash
# round-2 ninth GPU pass (1 GPU): marginals merged with the joint histogram (single-barrier finalize in every block), queue slots claimed with a shared atomic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest.log
P=gpurun_out/r2j_probe.jsonl; E=gpurun_out/r2j_probe.err; : > $P; : > $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 1,3 --stamps --tag scalar_fp32 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 1,3 --stamps --reps 3 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 1,3 --reps 2 --grid-poses 2048 >> $P 2>> $E
python tools/show_probe.py $P | cut -c1-300; tail -n 5 $E
