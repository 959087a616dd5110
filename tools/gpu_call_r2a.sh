#!/bin/bash
# round 2, first GPU call: validate the second-generation GEMM, then the whole GPU suite and a first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python tools/gemm_check.py > gpurun_out/r2a_gemm_check.log 2>&1
rc=$?
echo "gemm_check rc=$rc" | tee -a gpurun_out/r2a_gemm_check.log
tail -5 gpurun_out/r2a_gemm_check.log
if [ $rc -ne 0 ]; then
  echo "second-generation kernel failed its check: running the rest on the first-generation kernel" | tee gpurun_out/r2a_note.txt
  export GIB_TC_DEBUG=1
fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
echo "bench rc=$?"
cat gpurun_out/r2a_bench_c2.json | head -c 3000
