"""GPU diagnostic: per-tensor gradient errors of the CUDA path vs the fp32 oracle (and the oracle's own
fp32-vs-fp64 noise) on well-conditioned molecules.  usage: python tools/diag_parity.py MODEL"""
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, ".")
from graphinvent_b200 import synthetic as S
from oracle import mpnn_oracle as O
from tests.test_gpu_parity import _build, _step, _well_conditioned

model = sys.argv[1]
C = O.make_constants(model)
sd = O.init_state_dict(C, seed={"AttGGNN": 16, "EMN": 16}.get(model, 11))
n, e = S.random_graphs(1500, 13, 5, 3, seed=21, min_atoms=0)
n2, e2 = S.corner_case_graphs(13, 8)
nodes = torch.from_numpy(np.concatenate([n2, n])).float()
edges = torch.from_numpy(np.concatenate([e2, e])).float()
keep = _well_conditioned(C, sd, nodes, edges)
print(model, "well-conditioned:", keep.numel(), "with bonds", int((edges[keep].sum((1,2,3))>0).sum()))
nodes, edges = nodes[keep], edges[keep]
target = torch.from_numpy(S.random_targets(nodes.shape[0], 625, seed=5))
loss_ref, out_ref, g_ref = O.train_step_grads(sd, C, nodes, edges, target)
out, loss, grads = _step(_build(C, sd), nodes, edges, target)
print("logits max err", (out - out_ref).abs().max().item(), "loss", loss, float(loss_ref))
rows = []
for k, g in g_ref.items():
    d = (grads[k] - g)
    rows.append((d.abs().max().item() / max(g.abs().max().item(), 1e-12), d.norm().item() / max(g.norm().item(), 1e-12), k))
for r in sorted(rows, reverse=True)[:14]:
    print("  maxrel %.2e  l2rel %.2e  %s" % r)
print("  ... best:")
for r in sorted(rows)[:4]:
    print("  maxrel %.2e  l2rel %.2e  %s" % r)
# per-molecule check: which molecules carry the error (gradient of the logits sum wrt a bias, per molecule)
if len(sys.argv) > 2:
    net = _build(C, sd)
    bad = []
    for b in range(nodes.shape[0]):
        _, _, gb_ref = O.train_step_grads(sd, C, nodes[b:b + 1], edges[b:b + 1], target[b:b + 1])
        _, _, gb = _step(net, nodes[b:b + 1], edges[b:b + 1], target[b:b + 1])
        w = max((gb[k] - gb_ref[k]).abs().max().item() / max(gb_ref[k].abs().max().item(), 1e-12) for k in gb_ref)
        bad.append(w)
        print("   molecule", int(keep[b]), "atoms", int(nodes[b].sum() / 2), "bonds", int(edges[b].sum()), "worst rel %.2e" % w)
