"""GPU: tcgen05 weight-gradient GEMM (TN, MN-major operands) vs fp64 and vs the SIMT split-K kernel."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib
P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def dw(G, X, M, N, K, use_tc):
    lib.gib_set_tensor_cores(1 if use_tc else 0)
    dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), st()), "dw")
    return dW, db, sc

shapes = [(2048, 128, 128), (4100, 256, 256), (23808, 256, 256), (23808, 256, 128), (13312, 512, 512), (13312, 384, 128),
          (155648, 256, 256), (155648, 512, 512), (13312, 48, 512), (13312, 256, 144), (20000, 608, 512)]
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    torch.manual_seed(M + N)
    G = torch.randn(M, N, device="cuda"); X = torch.randn(M, K, device="cuda")
    ref = G.double().t() @ X.double(); refb = G.double().sum(0)
    out = {}
    for use_tc in (0, 1):
        dW, db, sc = dw(G, X, M, N, K, use_tc)
        torch.cuda.synchronize()
        e = (dW.double() - ref).abs().max().item(); eb = (db.double() - refb).abs().max().item()
        for _ in range(2): check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), st()), "dw")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), st()), "dw")
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 10
        out[use_tc] = (e, eb, t)
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:4d} K={K:4d}: err dW simt {out[0][0]:.2e} tc {out[1][0]:.2e} | db simt {out[0][1]:.2e} tc {out[1][1]:.2e} | "
          f"time simt {out[0][2]*1e3:7.1f} us ({fl/out[0][2]/1e9:5.1f} TF/s)  tc {out[1][2]*1e3:7.1f} us ({fl/out[1][2]/1e9:5.1f} TF/s)", flush=True)
lib.gib_set_tensor_cores(1)
