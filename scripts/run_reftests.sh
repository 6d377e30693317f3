#!/bin/bash
# Runs the prebuilt reference test programs (oracle/_ref/reftests, see oracle/Makefile reftests) on the
# GPU box and records exit codes + logs under gpurun_out/.  Usage: gpurun -- 'bash scripts/run_reftests.sh'
mkdir -p gpurun_out
: > gpurun_out/reftests_status.txt
for t in mesh-test band-matrix-and-solver-test bspline-test interpolation-test interpolation-template-test; do
  s=$SECONDS
  timeout 120 oracle/_ref/reftests/$t > gpurun_out/reftest_$t.log 2>&1
  rc=$?

  echo "$t exit $rc ($((SECONDS - s)) s)" >> gpurun_out/reftests_status.txt
done
cat gpurun_out/reftests_status.txt
