#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <script> [gpus] — retries while the pod answers "busy" (exit code 3)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpurun_last.log | tail -100; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
