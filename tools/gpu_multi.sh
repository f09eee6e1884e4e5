#!/bin/bash
# usage: tools/gpu_multi.sh <tag> <ngpus>   (run under gpurun --gpus N)
TAG=${1:-m01}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist check"; timeout 600 $RUN --master-port 29512 tools/dist_check.py > $OUT/${TAG}_dist_check.log 2>&1; echo "rc=$?"; grep -E "DIST_CHECK|P2P_CHECK|POSE_SHARD_CHECK|BFGS_SHARD_CHECK|Error|error" $OUT/${TAG}_dist_check.log | head
echo "== bench N=$N"; timeout 900 $RUN --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "rc=$?"; grep '^{' $OUT/${TAG}_bench_n$N.json; tail -3 $OUT/${TAG}_bench_n$N.err
[ "$QUICK" == "1" ] || { echo "== bench N=$N (nccl callback exchange)"; timeout 900 $RUN --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --exchange nccl > $OUT/${TAG}_bench_nccl_n$N.json 2> $OUT/${TAG}_bench_nccl_n$N.err; echo "rc=$?"; grep '^{' $OUT/${TAG}_bench_nccl_n$N.json | cut -c1-300; }
echo "== bench reference N=$N"; timeout 600 $RUN --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref_n$N.json 2>&1; echo "rc=$?"; grep '^{' $OUT/${TAG}_bench_ref_n$N.json | cut -c1-300
echo "== bench N=1 on the same box"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2>&1; grep '^{' $OUT/${TAG}_bench_n1.json | cut -c1-400
