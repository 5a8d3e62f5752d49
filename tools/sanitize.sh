#!/bin/bash
# compute-sanitizer memcheck + racecheck over tools/sanitize_case.py.
# usage: tools/sanitize.sh <tag>   -> gpurun_out/sanitize_<tool>_<tag>.log
TAG=${1:-x}
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 python tools/sanitize_case.py all > gpurun_out/sanitize_${TOOL}_$TAG.log 2>&1
  echo "$TOOL rc=$?" >> gpurun_out/sanitize_${TOOL}_$TAG.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|ok" gpurun_out/sanitize_${TOOL}_$TAG.log | tail -6
done
