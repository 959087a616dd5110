#!/bin/bash
# round 2, 1-GPU call: logit error of the shipped checkpoint under the split variants, full GPU tests, bench lines
mkdir -p gpurun_out
O=gpurun_out
for dbg in 0 4 $((512<<8)); do
  echo "== GIB_TC_DEBUG=$dbg"; GIB_TC_DEBUG=$dbg timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k pretrained 2>&1 | grep -E "pretrained/gdb13|passed|failed"
done > $O/r2m_logit_error.txt 2>&1; cat $O/r2m_logit_error.txt
timeout 300 python tools/gemm_check.py quick > $O/r2m_gemm_check.log 2>&1; echo "gemm_check rc=$?"; grep -E "GEMM_CHECK|155648" $O/r2m_gemm_check.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2m_pytest.log | cut -c1-300
timeout 300 python tools/tc3_probe.py 155648x256x256 23808x256x256 > $O/r2m_tc3_probe.log 2>&1; grep -E "==|product" $O/r2m_tc3_probe.log
timeout 600 python bench.py --steps 30 --warmup 5 --launch-table $O/r2m_launch_table_c2.txt > $O/r2m_bench_c2.json 2> $O/r2m_bench_c2.err; echo "bench rc=$?"; head -c 1200 $O/r2m_bench_c2.json; echo
timeout 600 python bench.py --steps 30 --warmup 5 --impl reference > $O/r2m_bench_c2_reference.json 2> $O/r2m_bench_c2_reference.err; echo "reference arm rc=$?"; head -c 800 $O/r2m_bench_c2_reference.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C4 --no-cpu-baseline --launch-table $O/r2m_launch_table_c4.txt > $O/r2m_bench_c4.json 2> $O/r2m_bench_c4.err; echo "bench C4 rc=$?"; head -c 1200 $O/r2m_bench_c4.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C3 --no-cpu-baseline > $O/r2m_bench_c3.json 2> $O/r2m_bench_c3.err; echo "bench C3 rc=$?"; head -c 1200 $O/r2m_bench_c3.json; echo
timeout 900 python tools/bench_generation.py --molecules 4000 > $O/r2m_generation_n1.json 2> $O/r2m_generation_n1.err; echo "generation rc=$?"; head -c 2500 $O/r2m_generation_n1.json; echo
