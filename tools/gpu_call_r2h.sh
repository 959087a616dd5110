#!/bin/bash
# round 2, 1-GPU call: N=256 MMA + pipelined splitters: check, probe, TN timeline, GPU tests, C2 / C4 bench lines
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gemm_check.py quick > $O/r2h_gemm_check.log 2>&1; rc=$?; echo "gemm_check rc=$rc"; grep -E "GEMM_CHECK|155648" $O/r2h_gemm_check.log | cut -c1-250
timeout 500 python tools/tc3_probe.py 155648x256x256 23808x256x256 > $O/r2h_tc3_probe.log 2>&1; echo "probe rc=$?"; grep -vE "MMA only|TMA only|interleaved|accumulators|round-to-nearest" $O/r2h_tc3_probe.log
timeout 120 python tools/tc3_trace.py 155648x256x256 6 tn > $O/r2h_tc3_trace.log 2>&1; cat $O/r2h_tc3_trace.log | cut -c1-200
if [ $rc -ne 0 ]; then export GIB_TC_DEBUG=$(( (32768|65536)<<8 )); echo "falling back"; fi
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2h_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r2h_bench_c2.json 2> $O/r2h_bench_c2.err; echo "bench rc=$?"; head -c 3000 $O/r2h_bench_c2.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C4 --no-cpu-baseline > $O/r2h_bench_c4.json 2> $O/r2h_bench_c4.err; echo "bench C4 rc=$?"; head -c 3000 $O/r2h_bench_c4.json; echo
