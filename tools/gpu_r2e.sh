#!/bin/bash
# round-2 two-GPU pass: in-kernel exchange checks (NCCL bootstrap), old dist checks, bench --gpus 2 (both configs, both arms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 tools/dist_check_pk.py > gpurun_out/r2e_dist_check_pk.log 2>&1; echo "dist_check_pk rc=$?"; grep PK_DIST_CHECK gpurun_out/r2e_dist_check_pk.log; tail -n 3 gpurun_out/r2e_dist_check_pk.log | cut -c1-200
timeout 900 $TR --master-port 29514 tools/dist_check.py > gpurun_out/r2e_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "CHECK" gpurun_out/r2e_dist_check.log | cut -c1-250; tail -n 3 gpurun_out/r2e_dist_check.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_peer or two_ranks" > gpurun_out/r2e_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2e_pytest_2gpu.log
timeout 900 $TR --master-port 29515 bench.py --gpus 2 --config C2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_c2_n2.json 2> gpurun_out/r2e_bench_c2_n2.err; echo "bench c2 n2 rc=$?"; cut -c1-500 gpurun_out/r2e_bench_c2_n2.json; tail -n 3 gpurun_out/r2e_bench_c2_n2.err | cut -c1-200
timeout 900 $TR --master-port 29516 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2e_bench_c3_n2.json 2> gpurun_out/r2e_bench_c3_n2.err; echo "bench c3 n2 rc=$?"; cut -c1-500 gpurun_out/r2e_bench_c3_n2.json; tail -n 3 gpurun_out/r2e_bench_c3_n2.err | cut -c1-200
timeout 600 $TR --master-port 29517 bench.py --gpus 2 --impl reference --config C2 --steps 2 --warmup 1 > gpurun_out/r2e_bench_c2_n2_ref.json 2> gpurun_out/r2e_bench_c2_n2_ref.err; echo "ref c2 n2 rc=$?"; cut -c1-300 gpurun_out/r2e_bench_c2_n2_ref.json
timeout 900 $TR --master-port 29518 bench.py --gpus 2 --config C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_c5_n2.json 2> gpurun_out/r2e_bench_c5_n2.err; echo "bench c5 n2 rc=$?"; cut -c1-500 gpurun_out/r2e_bench_c5_n2.json; tail -n 3 gpurun_out/r2e_bench_c5_n2.err | cut -c1-200
