#!/bin/bash
# Runs HERE: gpurun with retries while the pod answers "busy" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh <timeout-seconds> '<command>'   [GPUS=n in the environment for --gpus n]
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
