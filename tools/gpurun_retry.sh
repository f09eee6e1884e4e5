#!/bin/bash
# usage: [GPUS=n] tools/gpurun_retry.sh <timeout-seconds> <log-file> <command...>   -- retries while the pod answers busy / transient
TO=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "${GPUS:-1}" --timeout "$TO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" "$LOG" && ! grep -q "pytest rc\|rc=0" "$LOG"; then
    sleep 90
    continue
  fi
  break
done
echo "gpurun_retry finished after $i attempt(s), rc=$rc" >> "$LOG"
