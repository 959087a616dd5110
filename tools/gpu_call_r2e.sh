#!/bin/bash
# round 2, fifth GPU call: warp-level arrivals, full-line epilogue stores, dW sizing fix, chunk cap 4096
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gemm_check.py quick > $O/r2e_gemm_check.log 2>&1; echo "gemm_check rc=$?"; grep -E "^NT M= *(23808|155648) N= 256 K= 256|^TN M= *(23808|155648) N= 256 K= 256|GEMM_CHECK|probe" $O/r2e_gemm_check.log | cut -c1-330
timeout 300 python tools/tc3_probe.py 155648x256x256 23808x256x256 > $O/r2e_tc3_probe.log 2>&1; echo "probe rc=$?"; cat $O/r2e_tc3_probe.log
timeout 120 python tools/tc3_trace.py 155648x256x256 10 > $O/r2e_tc3_trace.log 2>&1; cat $O/r2e_tc3_trace.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $O/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $O/r2e_pytest.log | cut -c1-300
grep -h "^fp64-anchored \[" $O/r2e_pytest.log | cut -c1-560
timeout 500 python bench.py --steps 100 --warmup 5 > $O/r2e_bench_c2.json 2> $O/r2e_bench_c2.err
echo "bench C2 rc=$?"; head -c 700 $O/r2e_bench_c2.json; echo; tail -3 $O/r2e_bench_c2.err | cut -c1-300
GIB_TC_DEBUG=2 timeout 500 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-k2-in-model > $O/r2e_bench_c2_nochain.json 2> $O/r2e_bench_c2_nochain.err
echo "bench C2 (no chain launches) rc=$?"; head -c 400 $O/r2e_bench_c2_nochain.json; echo
timeout 500 python bench.py --config C4 --steps 30 --warmup 3 --no-k2-in-model > $O/r2e_bench_c4.json 2> $O/r2e_bench_c4.err
echo "bench C4 rc=$?"; head -c 400 $O/r2e_bench_c4.json; echo; tail -3 $O/r2e_bench_c4.err | cut -c1-300
