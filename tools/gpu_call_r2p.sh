#!/bin/bash
# round 2, last 1-GPU call: bench line with the tcgen05 / fp32-SIMT launch classes separated
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 --launch-table $O/r2p_launch_table_c2.txt > $O/r2p_bench_c2.json 2> $O/r2p_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench_c2.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'])
for k,v in d['roofline']['classes'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='note'})
PY
tail -3 $O/r2p_bench_c2.err
