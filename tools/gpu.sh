#!/bin/bash
# developer tool: run a command on the GPU box, retrying while the pod answers "busy" (exit code 3)
# usage: tools/gpu.sh <log name> <timeout s> [--gpus N] -- '<command>'
name=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$@" > "gpurun_out/call_$name.log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc attempt=$attempt"; tail -25 "gpurun_out/call_$name.log"; exit $rc; fi
  sleep 90
done
echo "gave up: pod busy"; exit 3
