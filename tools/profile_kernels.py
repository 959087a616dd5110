"""Launches the hot kernels at representative shapes so that `ncu` can capture them in isolation.

    python tools/profile_kernels.py scatter        # K2 at the C4 single-GPU shape, bond-type-grouped layout (204 MB > L2)
    python tools/profile_kernels.py gemm           # tcgen05 NT GEMM, C2 edge-MLP hidden layer (23808 x 256 x 256), planes
    python tools/profile_kernels.py dw             # tcgen05 weight-gradient GEMM of the same layer (+ its reduction)
    python tools/profile_kernels.py step C4 [B]    # one eager training step of a configuration (per-atom kernels:
                                                   # gru_*, graph_gather_*, k0_*, gather_rows, scatter_*)
usage on the GPU box, e.g.
    ncu --set full --clock-control none --import-source on -k regex:tc3_gemm -s 2 -c 2 -o gpurun_out/prof_gemm \
        python tools/profile_kernels.py gemm
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
what = sys.argv[1] if len(sys.argv) > 1 else "scatter"
reps = 4

if what == "scatter":
    S, E, ld = 155648, 352256, 112
    g = torch.Generator().manual_seed(0)
    dst = torch.randint(0, S, (E,), generator=g).sort().values
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    ptr[1:] = torch.bincount(dst, minlength=S).cumsum(0).int()
    t = torch.multinomial(torch.tensor([0.84, 0.14, 0.02]), E, replacement=True, generator=g)
    order = torch.argsort(t, stable=True)
    ent = torch.empty(E, dtype=torch.int32)
    ent[order] = torch.arange(E, dtype=torch.int32)          # message rows grouped by bond type, as in the model
    ptr, ent = ptr.cuda(), ent.cuda()
    msg, out = torch.randn(E, ld, device="cuda"), torch.empty(S, ld, device="cuda")
    for _ in range(reps):
        check(lib.gib_scatter_sum(P(out), P(msg), ld, P(ptr), P(ent), None, S, st()), "scatter")
elif what in ("gemm", "dw"):
    M, N, K = 23808, 256, 256                      # C2: single-bond entries x enn hidden layer
    if len(sys.argv) > 2:
        M, N, K = (int(v) for v in sys.argv[2].split("x"))
    X, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda")
    if what == "gemm":
        hi, lo = torch.empty_like(W), torch.empty_like(W)
        check(lib.gib_split_planes(P(W), P(hi), P(lo), W.numel(), st()), "split")
        Y = torch.empty(M, N, device="cuda")
        for _ in range(reps):
            check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, None, None, st()), "linear")
    else:
        G = torch.randn(M, N, device="cuda")
        dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
        sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
        for _ in range(reps):
            check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), None, None, st()), "dw")
elif what == "step":
    sys.path.insert(0, ".")
    import bench
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200.gnn import mpnn
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else None
    C, nodes, edges, target, _ = bench.make_batch(cfg, 1004, batch=B)
    torch.manual_seed(0)
    net = mpnn.create(C).cuda()
    nodes, edges, target = nodes.cuda(), edges.cuda(), target.cuda()
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        Fn.kl_loss(net(nodes, edges), target).backward()
    print("entries", net.last_stats.get("entries"), "workspace MB", net.last_stats.get("workspace_bytes", 0) / 1e6)
torch.cuda.synchronize()
print("done", what)
