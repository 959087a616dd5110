"""Launches the hot kernels at representative shapes so that `ncu` can capture them in isolation:
   K2 scatter-aggregate at the C4 single-GPU shape (working set 204 MB > L2) and the dense GEMMs at the
   C2 edge-MLP shape.  usage (on the GPU box):
     ncu --set full --clock-control none --import-source on -k regex:scatter_sum -c 2 -o gpurun_out/prof_scatter \
         python tools/profile_kernels.py scatter
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4

if what in ("scatter", "all"):
    S, E, ld = 155648, 352256, 112
    g = torch.Generator().manual_seed(0)
    dst = torch.randint(0, S, (E,), generator=g).sort().values
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    ptr[1:] = torch.bincount(dst, minlength=S).cumsum(0).int()
    ptr, ent = ptr.cuda(), torch.arange(E, dtype=torch.int32).cuda()
    msg, w, out = torch.randn(E, ld, device="cuda"), torch.ones(E, device="cuda"), torch.empty(S, ld, device="cuda")
    for _ in range(reps):
        check(lib.gib_scatter_sum(P(out), P(msg), ld, P(ptr), P(ent), P(w), S, st()), "scatter")
if what in ("gemm", "all"):
    M, N, K = 23808, 256, 256                      # C2: single-bond entries x enn hidden layer
    X, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda")
    for _ in range(reps):
        check(lib.gib_linear_fwd(P(X), K, P(W), K, P(b), P(Y), N, M, N, K, 1, st()), "linear")
    G = torch.randn(M, N, device="cuda")
    dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), st()), "dw")
torch.cuda.synchronize()
print("done", what)
