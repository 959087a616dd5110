#!/bin/bash
# round 2, last 2-GPU call: multi-GPU tests, generation tests (incl. the AttentionGGNN dummy-slot test), N=2 bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_generation.py tests/test_generation_rl.py tests/test_zz_rl_rollout_gpu.py -m gpu -q --timeout 500 > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2o_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 5 > $O/r2o_bench_c4_n2.json 2> $O/r2o_bench_c4_n2.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench_c4_n2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','dp_grad_rel_err','dp_grad_rel_err_fp32_gemms')}, d['single_gpu_same_workload']['value'], d['e2e']['value'])
PY
tail -3 $O/r2o_bench_c4_n2.err | cut -c1-300
