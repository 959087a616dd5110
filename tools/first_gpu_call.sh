#!/usr/bin/env bash
# First GPU call of the next round: regression + every diagnosis prepared without GPU time (tools/README.md).
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_pytest_tail.txt
python tools/tc_timing.py                         > gpurun_out/tc_timing.txt 2>&1
python tools/tc_timing.py --split2                > gpurun_out/tc_timing_split2.txt 2>&1
timeout 240 python tools/tc_variants.py 0 4 64 256 > gpurun_out/variants.txt 2>&1
timeout 150 python tools/tc2_test.py              > gpurun_out/tc2_test.txt 2>&1
timeout 150 python tools/tc_variants.py 128       > gpurun_out/variants_pair.txt 2>&1
timeout 150 python tools/tc_variants.py 192       > gpurun_out/variants_pair_raw.txt 2>&1
timeout 150 python tools/tc_timing.py --mask 128  > gpurun_out/tc_timing_pair.txt 2>&1
# ncu evidence for the HBM-bound per-atom kernels besides K2 (north_star: GRU update, graph gather, K0)
timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'gru_fwd_kernel|gru_bwd_kernel|graph_gather_fwd_kernel|graph_gather_bwd_kernel|k0_fill_kernel|gather_rows_kernel' \
  -c 12 -f -o gpurun_out/r2_per_atom_kernels python bench.py --config C4 --steps 1 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r2_ncu_per_atom.log 2>&1
for f in tc_timing tc_timing_split2 variants tc2_test variants_pair variants_pair_raw tc_timing_pair; do
  echo "==== $f"; tail -n 30 gpurun_out/$f.txt
done
