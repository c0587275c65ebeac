#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh [--gpus N] <timeout> '<command>'
GPUS=""
if [ "$1" == "--gpus" ]; then GPUS="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPUS --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
