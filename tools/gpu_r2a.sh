#!/bin/bash
# round-2 first GPU pass: full GPU test-suite, persistent-kernel probes (C2 / C3 / C5), one ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --timeout 600 -rf > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest.log
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps > gpurun_out/r2a_probe.jsonl 2> gpurun_out/r2a_probe.err
echo "probe c2 rc=$?"
VLCAL_PK_ATOM=1 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --tag atom1 >> gpurun_out/r2a_probe.jsonl 2>> gpurun_out/r2a_probe.err
VLCAL_PK_KPT=2 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --tag kpt2 >> gpurun_out/r2a_probe.jsonl 2>> gpurun_out/r2a_probe.err
VLCAL_PK_KPT=4 VLCAL_PK_ATOM=1 timeout 300 python tools/pk_probe.py --config C2 --modes 3 --tag kpt4_atom1 >> gpurun_out/r2a_probe.jsonl 2>> gpurun_out/r2a_probe.err
timeout 600 python tools/pk_probe.py --config C3 --modes 1,3 --stamps --reps 3 >> gpurun_out/r2a_probe.jsonl 2>> gpurun_out/r2a_probe.err
echo "probe c3 rc=$?"
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 1024 >> gpurun_out/r2a_probe.jsonl 2>> gpurun_out/r2a_probe.err
echo "probe c5 rc=$?"
cat gpurun_out/r2a_probe.jsonl
tail -20 gpurun_out/r2a_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2a_pk python tools/pk_probe.py --config C2 --modes 3 --reps 1 > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/r2a_ncu.log
