#!/bin/bash
# usage: scratch/gpucall.sh TAG TIMEOUT [--gpus N] -- 'command'   (retries while the pod answers "busy")
TAG=$1; shift
TMO=$1; shift
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift; shift; fi
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO $EXTRA -- "$@" > gpurun_out/${TAG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "attempt $attempt rc=$rc" >> gpurun_out/${TAG}_call.log; exit $rc; fi
  sleep 90
done
