#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <gpus> <script-under-repo>   (retries while the pod answers busy)
T=$1; N=$2; S=$3
for i in $(seq 1 40); do
  if [ "$N" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S"; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "bash $S"; rc=$?; fi
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
