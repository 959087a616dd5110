#!/bin/bash
# round 2, 2-GPU call: data-parallel tests, strong-scaling C4 bench at N=2, C5 generation with 2 replicas
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2g_smi.txt
timeout 300 python tools/gemm_check.py quick > $O/r2g_gemm_check.log 2>&1; rc=$?; echo "gemm_check rc=$rc"; grep -E "GEMM_CHECK|155648" $O/r2g_gemm_check.log | cut -c1-250
if [ $rc -ne 0 ]; then export GIB_TC_DEBUG=$((32768<<8)); echo "falling back to twelve MMAs per k-block"; fi
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $O/r2g_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -8 $O/r2g_pytest_multi.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > $O/r2g_bench_c4_n2.json 2> $O/r2g_bench_c4_n2.err
echo "bench N=2 rc=$?"; head -c 1500 $O/r2g_bench_c4_n2.json; echo; tail -5 $O/r2g_bench_c4_n2.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 --impl reference > $O/r2g_bench_c4_n2_reference.json 2> $O/r2g_bench_c4_n2_reference.err
echo "reference arm N=2 rc=$?"; head -c 600 $O/r2g_bench_c4_n2_reference.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/bench_generation.py --molecules 4000 > $O/r2g_generation_n2.json 2> $O/r2g_generation_n2.err
echo "generation N=2 rc=$?"; head -c 1500 $O/r2g_generation_n2.json; echo; tail -5 $O/r2g_generation_n2.err | cut -c1-300
