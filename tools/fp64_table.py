"""GPU diagnostic: gradient error of the CUDA path against the fp64 evaluation of the reference expression on an
ARBITRARY batch, beside the reference's own fp32 error -- tensor cores on (3xTF32) and off (fp32 SIMT GEMMs).
usage: python tools/fp64_table.py [MODEL ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from graphinvent_b200 import synthetic as S
from graphinvent_b200._lib import lib
from oracle import mpnn_oracle as O
from tests.test_gpu_parity import _build, _step

for model in (sys.argv[1:] or ["GGNN", "MNN", "AttGGNN", "EMN"]):
    C = O.make_constants(model)
    sd = O.init_state_dict(C, seed=11)
    n, e = S.random_graphs(96, 13, 5, 3, seed=12, min_atoms=0)
    n2, e2 = S.corner_case_graphs(13, 8)
    nodes = torch.from_numpy(np.concatenate([n2, n])).float()
    edges = torch.from_numpy(np.concatenate([e2, e])).float()
    target = torch.from_numpy(S.random_targets(nodes.shape[0], 625, seed=3))
    l32, o32, g32 = O.train_step_grads(sd, C, nodes, edges, target)
    l64, o64, g64 = O.train_step_grads(sd, C, nodes, edges, target, dtype=torch.float64)
    res = {}
    for tc in (1, 0):
        lib.gib_set_tensor_cores(tc)
        out, loss, grads = _step(_build(C, sd), nodes, edges, target)
        res[tc] = (out, loss, grads)
    lib.gib_set_tensor_cores(1)
    gl = lambda G: sum(((G[k].double() - g64[k]) ** 2).sum().item() for k in g64) ** 0.5
    gn = sum((g64[k] ** 2).sum().item() for k in g64) ** 0.5
    print(f"== {model}: |g| {gn:.3e}; global L2 err vs fp64: ref32 {gl(g32):.2e}  cuda-tc {gl(res[1][2]):.2e}  cuda-simt {gl(res[0][2]):.2e};"
          f" logits vs fp64: ref32 {(o32.double()-o64).abs().max():.2e} tc {(res[1][0].double()-o64).abs().max():.2e} simt {(res[0][0].double()-o64).abs().max():.2e}")
    rows = []
    for k, g in g64.items():
        nk = g.norm().item()
        rows.append(((res[1][2][k].double() - g).norm().item() / max(nk, 1e-30), (res[0][2][k].double() - g).norm().item() / max(nk, 1e-30),
                     (g32[k].double() - g).norm().item() / max(nk, 1e-30), nk, k))
    for r in sorted(rows, reverse=True)[:10]:
        print("   rel-L2 vs fp64: tc %.2e simt %.2e ref32 %.2e   |g| %.2e  %s" % r)
