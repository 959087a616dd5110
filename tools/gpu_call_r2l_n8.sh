#!/bin/bash
# round 2, 8-GPU call (charged 8x: keep it short): strong-scaling C4 bench at N=8, C5 generation with 8 replicas
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2l_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 30 --warmup 5 > $O/r2l_bench_c4_n8.json 2> $O/r2l_bench_c4_n8.err
echo "bench N=8 rc=$?"; head -c 1800 $O/r2l_bench_c4_n8.json; echo; grep -v "^$\|OMP_NUM\|\*\*\*" $O/r2l_bench_c4_n8.err | tail -5 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/bench_generation.py --molecules 12500 --no-cpu > $O/r2l_generation_n8.json 2> $O/r2l_generation_n8.err
echo "generation N=8 rc=$?"; head -c 1500 $O/r2l_generation_n8.json; echo; grep -v "it/s\|^$\|OMP_NUM\|\*\*\*" $O/r2l_generation_n8.err | tail -5 | cut -c1-300
