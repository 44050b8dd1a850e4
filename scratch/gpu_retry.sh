#!/bin/bash
# usage: scratch/gpu_retry.sh <timeout-seconds> <gpus> <command...> ; retries while the pod answers busy
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -80; exit 0
done
echo "gave up: pod busy"; exit 3
