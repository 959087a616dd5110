#!/bin/bash
# round 2, sixth GPU call: MMA-rate and epilogue experiments (probe only) + timeline
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/tc3_probe.py 155648x256x256 23808x256x256 > $O/r2f_tc3_probe.log 2>&1; echo "probe rc=$?"; cat $O/r2f_tc3_probe.log
timeout 120 python tools/tc3_trace.py 155648x256x256 6 > $O/r2f_tc3_trace.log 2>&1; head -12 $O/r2f_tc3_trace.log
