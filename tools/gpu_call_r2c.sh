#!/bin/bash
# round 2, third GPU call: compact epilogues -- correctness, probe, full suite, bench lines
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gemm_check.py quick > $O/r2c_gemm_check.log 2>&1; echo "gemm_check rc=$?"; tail -12 $O/r2c_gemm_check.log | cut -c1-250
timeout 300 python tools/tc3_probe.py 155648x256x256 23808x256x256 23808x256x128 > $O/r2c_tc3_probe.log 2>&1; echo "probe rc=$?"; cat $O/r2c_tc3_probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 $O/r2c_pytest.log | cut -c1-300
grep -h "^fp64-anchored \[" $O/r2c_pytest.log | cut -c1-400
timeout 500 python bench.py --steps 100 --warmup 5 > $O/r2c_bench_c2.json 2> $O/r2c_bench_c2.err
echo "bench C2 rc=$?"; head -c 1500 $O/r2c_bench_c2.json; echo; tail -3 $O/r2c_bench_c2.err | cut -c1-300
timeout 500 python bench.py --config C4 --steps 30 --warmup 3 --no-k2-in-model > $O/r2c_bench_c4.json 2> $O/r2c_bench_c4.err
echo "bench C4 rc=$?"; head -c 700 $O/r2c_bench_c4.json; echo; tail -3 $O/r2c_bench_c4.err | cut -c1-300
timeout 500 python bench.py --config C3 --steps 20 --warmup 3 --no-k2-in-model > $O/r2c_bench_c3.json 2> $O/r2c_bench_c3.err
echo "bench C3 rc=$?"; head -c 700 $O/r2c_bench_c3.json; echo; tail -3 $O/r2c_bench_c3.err | cut -c1-300
