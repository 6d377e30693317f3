#!/bin/bash
# Per-sweep times of the 512^3 solve (axis 1 and axis 0 sweeps) against the resident warps per SM of the
# L2-resident sweep.  Usage: scripts/sweep_l2_scan.sh [periodic]
run() { python scripts/sweep_times.py 512 ${PER:-} 2>&1 | sed -n 2,4p | tr '\n' ' '; echo; }
PER=$1
export BSPL_SWEEP_L2=0; echo -n "thread-per-line: "; run
export BSPL_SWEEP_L2=1
for w in ${WARPS:-3 4 5 6}; do
  export BSPL_SWEEP_L2_WARPS=$w
  echo -n "warps $w: "; run
done
