#!/bin/bash
# round 2, final 1-GPU confirmation: full GPU test suite, smoke, default bench line (+ launch table)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2n_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > $O/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2n_smoke.log | cut -c1-300
timeout 600 python bench.py --launch-table $O/r2n_launch_table_c2.txt > $O/r2n_bench_c2.json 2> $O/r2n_bench_c2.err; echo "bench rc=$?"; head -c 1500 $O/r2n_bench_c2.json; echo
GIB_TC_DEBUG=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k pretrained 2>&1 | grep -E "pretrained/gdb13|passed|failed"
