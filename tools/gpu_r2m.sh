#!/bin/bash
# round-2 evidence pass (1 GPU) for the final kernel: suite, smoke, probes (packed vs scalar fp32 A/B), both bench arms, ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 -rf > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2m_smoke.log
P=gpurun_out/r2m_probe.jsonl; E=gpurun_out/r2m_probe.err; : > $P; : > $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag scalar_fp32 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 1,3 --stamps --reps 3 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 3 --tag scalar_fp32 >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 >> $P 2>> $E
VLCAL_LIB=$ALT timeout 600 python tools/pk_probe.py --config C5 --modes 3 --reps 2 --grid-poses 2048 --tag scalar_fp32 >> $P 2>> $E
python tools/show_probe.py $P | cut -c1-300; tail -n 3 $E
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
echo "bench c3 rc=$?"; cut -c1-200 gpurun_out/r2m_bench_c3.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_c3_ref.json 2> gpurun_out/r2m_bench_c3_ref.err
echo "ref c3 rc=$?"; cut -c1-200 gpurun_out/r2m_bench_c3_ref.json
timeout 600 python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err
echo "bench c2 rc=$?"; cut -c1-200 gpurun_out/r2m_bench_c2.json
timeout 900 python bench.py --config C5 --steps 2 --warmup 3 > gpurun_out/r2m_bench_c5.json 2> gpurun_out/r2m_bench_c5.err
echo "bench c5 rc=$?"; cut -c1-200 gpurun_out/r2m_bench_c5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_launches_bench.log 2>&1
echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2m_pk_c2 python tools/pk_probe.py --config C2 --modes 3 --reps 1 > gpurun_out/r2m_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nid_persistent -c 1 -o gpurun_out/r2m_pk_c3 python tools/pk_probe.py --config C3 --modes 3 --reps 1 > gpurun_out/r2m_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
ls -la gpurun_out/r2m*.ncu-rep gpurun_out/r2m_launches.csv
