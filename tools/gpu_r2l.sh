#!/bin/bash
# diagnostic pass (1 GPU): finer stamps inside the finalizer (alt build), upload team size A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
P=gpurun_out/r2l_probe.jsonl; E=gpurun_out/r2l_probe.err; : > $P; : > $E
ALT=$PWD/direct_visual_lidar_calibration_b200/libvlcal_nid_alt.so
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C2 --modes 3 --stamps --tag diag_stamps >> $P 2>> $E
VLCAL_LIB=$ALT timeout 300 python tools/pk_probe.py --config C3 --modes 3 --stamps --reps 2 --tag diag_stamps >> $P 2>> $E
python tools/show_probe.py $P | cut -c1-300; tail -n 3 $E
for th in 8 32 48; do
VLCAL_UPLOAD_THREADS=$th timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_c3_up$th.json 2> gpurun_out/r2l_bench_c3_up$th.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_bench_c3_up$th.json").read().strip().splitlines()[-1])
print("upload threads $th:", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["host_breakdown_ms_per_solve"])
PY
done
VLCAL_UPLOAD_THREADS=32 timeout 600 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_c2_up32.json 2> gpurun_out/r2l_bench_c2_up32.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_bench_c2_up32.json").read().strip().splitlines()[-1])
print("C2 upload threads 32:", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["host_breakdown_ms_per_solve"])
PY
