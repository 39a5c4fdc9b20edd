#!/bin/bash
# gpurun with retry while the pod's GPU slots are busy (exit code 3 = nothing charged).
# usage: tools/grun.sh <timeout_s> <gpus> '<command>'
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
