#!/bin/bash
# round 2, 1-GPU call: branch-free SELU epilogue + aux prefetch: check, probe, kernel tests, C2 / C4 bench lines + launch tables
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/gemm_check.py quick > $O/r2j_gemm_check.log 2>&1; rc=$?; echo "gemm_check rc=$rc"; grep -E "GEMM_CHECK|155648" $O/r2j_gemm_check.log | cut -c1-250
timeout 500 python tools/tc3_probe.py 155648x256x256 23808x256x256 23808x256x128 > $O/r2j_tc3_probe.log 2>&1; echo "probe rc=$?"; grep -vE "MMA only|TMA only|interleaved|accumulators|round-to-nearest" $O/r2j_tc3_probe.log
timeout 120 python tools/tc3_trace.py 155648x256x256 6 > $O/r2j_tc3_trace.log 2>&1; head -9 $O/r2j_tc3_trace.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_capacity.py -m gpu -x -q --timeout 900 > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2j_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --launch-table $O/r2j_launch_table_c2.txt > $O/r2j_bench_c2.json 2> $O/r2j_bench_c2.err; echo "bench rc=$?"; head -c 3000 $O/r2j_bench_c2.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C4 --no-cpu-baseline --launch-table $O/r2j_launch_table_c4.txt > $O/r2j_bench_c4.json 2> $O/r2j_bench_c4.err; echo "bench C4 rc=$?"; head -c 3000 $O/r2j_bench_c4.json; echo
