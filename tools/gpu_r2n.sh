#!/bin/bash
# round-2 eight-GPU pass: in-kernel exchange at world 8, bench --gpus 8 / 4 (C3 default, C2), reference arm at N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR8 --master-port 29613 tools/dist_check_pk.py > gpurun_out/r2n_dist_check_pk_8gpu.log 2>&1; echo "dist_check_pk rc=$?"; grep PK_DIST_CHECK gpurun_out/r2n_dist_check_pk_8gpu.log | cut -c1-250; tail -n 2 gpurun_out/r2n_dist_check_pk_8gpu.log | cut -c1-200
timeout 900 $TR8 --master-port 29614 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2n_bench_c3_n8.json 2> gpurun_out/r2n_bench_c3_n8.err; echo "bench c3 n8 rc=$?"; tail -n 2 gpurun_out/r2n_bench_c3_n8.err | cut -c1-200
timeout 900 $TR8 --master-port 29615 bench.py --gpus 8 --config C2 --steps 5 --warmup 3 > gpurun_out/r2n_bench_c2_n8.json 2> gpurun_out/r2n_bench_c2_n8.err; echo "bench c2 n8 rc=$?"; tail -n 2 gpurun_out/r2n_bench_c2_n8.err | cut -c1-200
timeout 900 $TR4 --master-port 29616 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2n_bench_c3_n4.json 2> gpurun_out/r2n_bench_c3_n4.err; echo "bench c3 n4 rc=$?"; tail -n 2 gpurun_out/r2n_bench_c3_n4.err | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c3_n1.json 2> gpurun_out/r2n_bench_c3_n1.err; echo "bench c3 n1 rc=$?"
timeout 900 $TR8 --master-port 29617 bench.py --gpus 8 --config C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_c5_n8.json 2> gpurun_out/r2n_bench_c5_n8.err; echo "bench c5 n8 rc=$?"; tail -n 2 gpurun_out/r2n_bench_c5_n8.err | cut -c1-200
timeout 600 $TR8 --master-port 29618 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_bench_c3_n8_ref.json 2> gpurun_out/r2n_bench_c3_n8_ref.err; echo "ref c3 n8 rc=$?"; cut -c1-300 gpurun_out/r2n_bench_c3_n8_ref.json
python - <<'PY'
import json
for f in ("r2n_bench_c3_n1","r2n_bench_c3_n4","r2n_bench_c3_n8","r2n_bench_c2_n8","r2n_bench_c5_n8"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["host_breakdown_ms_per_solve"], d["run"].get("host_cores_of_rank0"), d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
