#!/bin/bash
# Runs the prebuilt reference test programs (oracle/_ref/reftests, see oracle/Makefile reftests) on the
# GPU box and records exit codes + logs under gpurun_out/.  Usage: gpurun -- 'bash scripts/run_reftests.sh'
mkdir -p gpurun_out
: > gpurun_out/reftests_status.txt
for t in mesh-test band-matrix-and-solver-test bspline-test interpolation-test interpolation-test.dummy-point interpolation-template-test; do
  s=$SECONDS
  timeout 120 oracle/_ref/reftests/$t > gpurun_out/reftest_$t.log 2>&1
  rc=$?

  echo "$t exit $rc ($((SECONDS - s)) s)" >> gpurun_out/reftests_status.txt
done
cat gpurun_out/reftests_status.txt
# the same workload as the reference's interpolation-speed-test, batched (scripts/cpp/speed_test_batched.cpp)
CXX=/usr/bin/g++; [ -x $CXX ] || CXX=g++
$CXX -std=c++17 -O2 -I include scripts/cpp/speed_test_batched.cpp -L bsplineinterpolation_b200 -lbspline_b200 \
  -Wl,-rpath,$PWD/bsplineinterpolation_b200 -o /tmp/speed_test_batched \
  && timeout 120 /tmp/speed_test_batched > gpurun_out/speed_test_batched.txt 2>&1
tail -22 gpurun_out/speed_test_batched.txt
