#!/bin/bash
# round 2, 1-GPU call: full GPU test suite, C2/C3/C4 bench lines + launch tables, ncu captures (per-atom kernels at C4,
# tcgen05 NT / TN, K2), launch list of a C2 bench run
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2k_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 --launch-table $O/r2k_launch_table_c2.txt > $O/r2k_bench_c2.json 2> $O/r2k_bench_c2.err; echo "bench rc=$?"; head -c 2500 $O/r2k_bench_c2.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C4 --no-cpu-baseline --launch-table $O/r2k_launch_table_c4.txt > $O/r2k_bench_c4.json 2> $O/r2k_bench_c4.err; echo "bench C4 rc=$?"; head -c 1500 $O/r2k_bench_c4.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --config C3 --no-cpu-baseline > $O/r2k_bench_c3.json 2> $O/r2k_bench_c3.err; echo "bench C3 rc=$?"; head -c 1500 $O/r2k_bench_c3.json; echo
# ncu: per-atom kernels of one eager C4 step (second step of the script), then the GEMMs and K2 in isolation
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gru_|graph_gather|k0_|gather_rows|scatter_sum|softmax|kl_loss|adam' -c 70 -o $O/r2k_step_c4 python tools/profile_kernels.py step C4 2048 > $O/r2k_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2k_tc3_nt python tools/profile_kernels.py gemm > $O/r2k_ncu_nt.log 2>&1; echo "ncu nt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 1 -o $O/r2k_tc3_tn python tools/profile_kernels.py dw > $O/r2k_ncu_tn.log 2>&1; echo "ncu tn rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:scatter_sum -s 2 -c 1 -o $O/r2k_scatter python tools/profile_kernels.py scatter > $O/r2k_ncu_scatter.log 2>&1; echo "ncu scatter rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2k_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-module-api --no-k2-in-model > $O/r2k_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
ls -la $O | grep r2k | awk '{print $5, $9}'
