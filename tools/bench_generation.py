"""GPU: molecules/s of batched generation (model forward + device-side round) with the shipped GGNN checkpoint.
usage: python tools/bench_generation.py [batch] [repeats]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from graphinvent_b200.config import make_constants
from graphinvent_b200.generation import GraphGenerator
from graphinvent_b200.gnn import mpnn
from tests.conftest import pretrained_path

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
C = make_constants("GGNN")
net = mpnn.create(C)
path = pretrained_path()
if path:
    net.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
net = net.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(0)
res = []
for r in range(reps + 1):
    gen = GraphGenerator(net, batch_size=batch, n_atom_types=5, n_formal_charge=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = gen.build_graphs(generator=g)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if r:
        res.append((n / dt, gen.rounds, n))
best = max(res)
(nodes, _, n_nodes) = (gen.generated_nodes[:batch], None, gen.generated_n_nodes[:batch])
print(json.dumps({"metric": "generated molecules/s (GGNN pretrained, device-side rounds)", "batch": batch,
                  "value": best[0], "rounds": best[1], "n_generated": best[2],
                  "mean_atoms": float(n_nodes.float().mean()), "properly_terminated": float(gen.properly_terminated[:batch].float().mean()),
                  "all_runs": [round(v[0]) for v in res], "checkpoint": bool(path)}))
