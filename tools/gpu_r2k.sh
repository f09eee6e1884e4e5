#!/bin/bash
# round-2 second two-GPU pass: the every-block finalize with the in-kernel exchange; bench --gpus 2 with per-rank core binding;
# single-GPU probe of the same build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 tools/dist_check_pk.py > gpurun_out/r2k_dist_check_pk.log 2>&1; echo "dist_check_pk rc=$?"; grep PK_DIST_CHECK gpurun_out/r2k_dist_check_pk.log; tail -n 3 gpurun_out/r2k_dist_check_pk.log | cut -c1-200
timeout 900 $TR --master-port 29514 tools/dist_check.py > gpurun_out/r2k_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "CHECK" gpurun_out/r2k_dist_check.log | cut -c1-250; tail -n 3 gpurun_out/r2k_dist_check.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_peer or two_ranks or persistent" > gpurun_out/r2k_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2k_pytest_2gpu.log
P=gpurun_out/r2k_probe.jsonl; E=gpurun_out/r2k_probe.err; : > $P; : > $E
timeout 400 python tools/pk_probe.py --config C2 --modes 1,3 --stamps >> $P 2>> $E
timeout 600 python tools/pk_probe.py --config C3 --modes 1,3 --stamps --reps 3 >> $P 2>> $E
python tools/show_probe.py $P | cut -c1-300; tail -n 3 $E
timeout 900 $TR --master-port 29516 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_c3_n2.json 2> gpurun_out/r2k_bench_c3_n2.err; echo "bench c3 n2 rc=$?"; cut -c1-300 gpurun_out/r2k_bench_c3_n2.json; tail -n 3 gpurun_out/r2k_bench_c3_n2.err | cut -c1-200
timeout 900 $TR --master-port 29515 bench.py --gpus 2 --config C2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_c2_n2.json 2> gpurun_out/r2k_bench_c2_n2.err; echo "bench c2 n2 rc=$?"; cut -c1-300 gpurun_out/r2k_bench_c2_n2.json; tail -n 3 gpurun_out/r2k_bench_c2_n2.err | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_c3_n1.json 2> gpurun_out/r2k_bench_c3_n1.err; echo "bench c3 n1 rc=$?"; cut -c1-300 gpurun_out/r2k_bench_c3_n1.json
python - <<'PY'
import json
for f in ("r2k_bench_c3_n1","r2k_bench_c3_n2","r2k_bench_c2_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["host_breakdown_ms_per_solve"], d["run"].get("host_cores_of_rank0"))
    except Exception as e:
        print(f, "unreadable", e)
PY
