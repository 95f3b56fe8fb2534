#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit code 3: nothing charged).
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i"; sleep 90
done
exit 3
