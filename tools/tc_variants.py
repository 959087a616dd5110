"""GPU: accuracy and speed of the tcgen05 GEMM variants behind gib_tc_debug, on the whole model.

    timeout 300 python tools/tc_variants.py            # every variant
    timeout 300 python tools/tc_variants.py 0 64       # a subset (gib_tc_debug masks)

Variants (include/gib200.h): 0 product (round-to-nearest hi/lo split), 4 truncation split, 64 raw hi operand
(SPLIT = 2), 256 product arithmetic with explicit LDS / STS (SPLIT = 3), 128 CTA-pair kernel (gemm_tc2.cu, cta_group::2 -- first run: wrap in `timeout`, its waits trap after
~10 s if a hand-off is wrong), 192 = 128 | 64 (pair kernel with the raw hi operand).
For each: max |logit - reference| on the shipped checkpoint x 256 real gdb13 rows (bonded / bond-less, bar 1e-4),
gradient norm-level deviation, and the C2 training-step time.
"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphinvent_b200 import functional as Fn  # noqa: E402
from graphinvent_b200._lib import lib  # noqa: E402
from graphinvent_b200.gnn import mpnn  # noqa: E402
from graphinvent_b200.optim import FlatAdam  # noqa: E402
from oracle import mpnn_oracle as O  # noqa: E402
from tests.conftest import load_gdb13, pretrained_path  # noqa: E402

dev = torch.device("cuda", 0)
modes = [int(a) for a in sys.argv[1:]] or [0, 4, 64, 256, 128, 192]


def ev(fn, K=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K


fx = load_gdb13()
path = pretrained_path()
pre = None
if path is not None:
    pre = mpnn.create(O.make_constants("GGNN"))
    pre.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
    pre = pre.to(dev)
bonded = fx["edges"].sum((1, 2, 3)) > 0
C, nodes_h, edges_h, target_h, apd = bench.make_batch("C2", 1002)
nodes, edges, target = nodes_h.to(dev), edges_h.to(dev), target_h.to(dev)
base_grads = None
for mode in modes:
    lib.gib_tc_debug(mode)
    line = f"debug {mode:3d}:"
    if pre is not None:
        pre.zero_grad()
        out = pre(fx["nodes"].to(dev), fx["edges"].to(dev))
        Fn.kl_loss(out, fx["apds"].to(dev)).backward()
        err = (out.detach().cpu() - fx["logits"]).abs().max(1).values
        grads = torch.cat([p.grad.flatten() for p in pre.parameters()]).clone()
        if base_grads is None:
            base_grads = grads
        line += (f" logits bonded max {err[bonded].max():.2e} mean {err[bonded].mean():.2e} bond-less max "
                 f"{err[~bonded].max() if (~bonded).any() else 0:.2e}; grads vs first variant rel-L2 "
                 f"{((grads - base_grads).norm() / base_grads.norm()).item():.2e};")
    torch.manual_seed(0)
    net = mpnn.create(C).to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-4)

    def step():
        loss = Fn.kl_loss(net(nodes, edges), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    ms = ev(step)
    line += f" C2 step {ms:.3f} ms = {nodes.shape[0] / ms * 1e3:,.0f} graphs/s"
    print(line, flush=True)
lib.gib_tc_debug(0)
