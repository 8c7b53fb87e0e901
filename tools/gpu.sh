#!/bin/bash
# retry wrapper around gpurun: tools/gpu.sh <timeout-seconds> '<command>'  (retries while the pod answers busy)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | grep -q "status=transient" || { echo "$out"; exit 0; }
  sleep 45
done
echo "$out"; exit 3
