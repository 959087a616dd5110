#!/bin/bash
# round 2, fourth GPU call: fast activation split + dependent-chain launches: correctness, probe, suite, benches, ncu
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/gemm_check.py quick > $O/r2d_gemm_check.log 2>&1; echo "gemm_check rc=$?"; grep -E "^NT M= *(23808|13312) N= 256|^TN M= *23808 N= 256 K= 256|GEMM_CHECK" $O/r2d_gemm_check.log | cut -c1-330
timeout 300 python tools/tc3_probe.py 155648x256x256 23808x256x256 > $O/r2d_tc3_probe.log 2>&1; echo "probe rc=$?"; cat $O/r2d_tc3_probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $O/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $O/r2d_pytest.log | cut -c1-300
grep -h "^fp64-anchored \[" $O/r2d_pytest.log | cut -c1-520
timeout 500 python bench.py --steps 100 --warmup 5 > $O/r2d_bench_c2.json 2> $O/r2d_bench_c2.err
echo "bench C2 rc=$?"; head -c 700 $O/r2d_bench_c2.json; echo; tail -3 $O/r2d_bench_c2.err | cut -c1-300
timeout 500 python bench.py --config C4 --steps 30 --warmup 3 --no-k2-in-model > $O/r2d_bench_c4.json 2> $O/r2d_bench_c4.err
echo "bench C4 rc=$?"; head -c 400 $O/r2d_bench_c4.json; echo; tail -3 $O/r2d_bench_c4.err | cut -c1-300
timeout 500 python bench.py --config C3 --steps 20 --warmup 3 --no-k2-in-model > $O/r2d_bench_c3.json 2> $O/r2d_bench_c3.err
echo "bench C3 rc=$?"; head -c 400 $O/r2d_bench_c3.json; echo; tail -3 $O/r2d_bench_c3.err | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2d_tc3_nt \
    python tools/profile_kernels.py gemm > $O/r2d_ncu_nt.log 2>&1; echo "ncu nt rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2d_tc3_tn \
    python tools/profile_kernels.py dw > $O/r2d_ncu_tn.log 2>&1; echo "ncu tn rc=$?"
