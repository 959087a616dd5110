#!/bin/bash
# round 2, first GPU call: validate the second-generation GEMM, then the GPU suite, bench lines and profiles
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 600 python tools/gemm_check.py > $O/r2a_gemm_check.log 2>&1
rc=$?
echo "gemm_check rc=$rc" | tee -a $O/r2a_gemm_check.log
tail -4 $O/r2a_gemm_check.log
if [ $rc -ne 0 ]; then
  echo "second-generation kernel failed its check: GPU suite on the first-generation kernel, then stop" | tee $O/r2a_note.txt
  GIB_TC_DEBUG=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/r2a_pytest_gen1.log 2>&1
  tail -15 $O/r2a_pytest_gen1.log
  exit 0
fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $O/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r2a_pytest.log
tail -25 $O/r2a_pytest.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r2a_bench_c2.json 2> $O/r2a_bench_c2.err
echo "bench C2 rc=$?"; head -c 1500 $O/r2a_bench_c2.json; echo
timeout 300 python tools/k2_variants.py > $O/r2a_k2_variants.log 2>&1; tail -1 $O/r2a_k2_variants.log | head -c 1200; echo
timeout 600 python bench.py --config C4 --steps 30 --warmup 3 --no-k2-in-model > $O/r2a_bench_c4.json 2> $O/r2a_bench_c4.err
echo "bench C4 rc=$?"; head -c 600 $O/r2a_bench_c4.json; echo
timeout 600 python bench.py --config C3 --steps 20 --warmup 3 --no-k2-in-model > $O/r2a_bench_c3.json 2> $O/r2a_bench_c3.err
echo "bench C3 rc=$?"; head -c 600 $O/r2a_bench_c3.json; echo
# launch list of the C2 step (eager pass: graph replays hide the kernels from ncu's per-launch list)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file $O/r2a_launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-k2-in-model > $O/r2a_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 2 -o $O/r2a_tc3_nt \
    python tools/profile_kernels.py gemm > $O/r2a_ncu_nt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 2 -o $O/r2a_tc3_tn \
    python tools/profile_kernels.py dw > $O/r2a_ncu_tn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scatter_sum -s 2 -c 2 -o $O/r2a_scatter \
    python tools/profile_kernels.py scatter > $O/r2a_ncu_scatter.log 2>&1
echo "ncu done"; ls -la $O | tail -20
