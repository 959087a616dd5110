#!/bin/bash
# round 2, second GPU call: where does the tcgen05 GEMM lose its time (probe + ncu), parity tables, full suite, bench
mkdir -p gpurun_out
O=gpurun_out
timeout 420 python tools/tc3_probe.py > $O/r2b_tc3_probe.log 2>&1; echo "probe rc=$?"; tail -45 $O/r2b_tc3_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2b_tc3_nt_big \
    python tools/profile_kernels.py gemm 155648x256x256 > $O/r2b_ncu_nt_big.log 2>&1; echo "ncu nt big rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2b_tc3_nt \
    python tools/profile_kernels.py gemm > $O/r2b_ncu_nt.log 2>&1; echo "ncu nt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2b_tc3_tn \
    python tools/profile_kernels.py dw > $O/r2b_ncu_tn.log 2>&1; echo "ncu tn rc=$?"
timeout 300 python tools/fp64_table.py > $O/r2b_fp64_table.log 2>&1; echo "fp64 table rc=$?"; cat $O/r2b_fp64_table.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_fp64_anchored_default_dims_arbitrary_batch \
    --deselect tests/test_gpu_parity.py::test_fp64_anchored_c2_slice --deselect tests/test_gpu_parity.py::test_fp64_anchored_pretrained_on_all_real_gdb13_rows \
    > $O/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $O/r2b_pytest.log | cut -c1-300
timeout 500 python bench.py --steps 100 --warmup 5 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err
echo "bench C2 rc=$?"; head -c 1800 $O/r2b_bench_c2.json; echo; tail -5 $O/r2b_bench_c2.err | cut -c1-300
timeout 200 python tools/k2_variants.py > $O/r2b_k2_variants.log 2>&1; tail -1 $O/r2b_k2_variants.log | head -c 1500; echo
