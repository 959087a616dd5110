"""GPU timing probe of the tcgen05 GEMM with parts of the pipeline switched off (results invalid in those modes)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = [(155648, 256, 256), (13312, 512, 512), (23808, 256, 256)]
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in sys.argv[1].split("x"))]
modes = [0, 4, 1, 2, 3] if len(sys.argv) <= 2 else [int(sys.argv[2])]
for (M, N, K) in shapes:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda")
    for mode in modes:
        for act in (0, 1):
            lib.gib_tc_debug(mode)
            for _ in range(3): check(lib.gib_linear_fwd_tc(P(X), K, P(W), K, P(b), P(Y), N, M, N, K, act, st()), "tc")
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): check(lib.gib_linear_fwd_tc(P(X), K, P(W), K, P(b), P(Y), N, M, N, K, act, st()), "tc")
            e.record(); torch.cuda.synchronize()
            t = a.elapsed_time(e) / 10
            print(f"M={M} N={N} K={K} mode={mode} act={act}: {t*1e3:.1f} us  {2.0*M*N*K/t/1e9:.1f} TF/s", flush=True)
lib.gib_tc_debug(0)
