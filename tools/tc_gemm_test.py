"""GPU: tcgen05 3xTF32 GEMM vs fp64 reference and vs the SIMT fp32 kernel (accuracy + speed)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(fn, X, W, b, Y, M, N, K, act):
    check(fn(P(X), K, P(W), K, P(b), P(Y), N, M, N, K, act, st()), "linear")


def bench(fn, args, iters=20):
    for _ in range(3):
        run(fn, *args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run(fn, *args)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


shapes = [(128, 128, 32), (100, 48, 16), (1000, 250 + 6, 256), (1000, 256, 256), (23808, 256, 256), (23808, 128, 256),
          (23808, 256, 128), (13312, 512, 512), (13312, 384, 128), (155648, 256, 256), (155648, 512, 512),
          (4096, 48, 512), (5000, 608, 512), (13312, 256, 144)]
only = sys.argv[1:] and [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in (only or shapes):
    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    for act in (0, 1):
        r = torch.selu(ref) if act else ref
        Ytc = torch.full((M, N), float("nan"), device="cuda")
        Ysi = torch.full((M, N), float("nan"), device="cuda")
        lib.gib_set_tensor_cores(0)
        run(lib.gib_linear_fwd, X, W, b, Ysi, M, N, K, act)
        try:
            run(lib.gib_linear_fwd_tc, X, W, b, Ytc, M, N, K, act)
            torch.cuda.synchronize()
        except Exception as ex:
            print(f"M={M} N={N} K={K} act={act}: TC FAILED {ex}")
            raise
        lib.gib_tc_debug(4)
        Yrn = torch.full((M, N), float("nan"), device="cuda")
        run(lib.gib_linear_fwd_tc, X, W, b, Yrn, M, N, K, act)
        lib.gib_tc_debug(0)
        e_rn = (Yrn.double() - r).abs().max().item()
        e_tc = (Ytc.double() - r).abs().max().item()
        e_si = (Ysi.double() - r).abs().max().item()
        bad = int((~torch.isfinite(Ytc)).sum())
        print(f"M={M:6d} N={N:4d} K={K:4d} act={act}: max err tc(trunc) {e_tc:.2e} tc(rna) {e_rn:.2e} simt {e_si:.2e}  nonfinite {bad}", flush=True)
    t_tc = bench(lib.gib_linear_fwd_tc, (X, W, b, Ytc, M, N, K, 1))
    t_si = bench(lib.gib_linear_fwd, (X, W, b, Ysi, M, N, K, 1))
    fl = 2.0 * M * N * K
    print(f"        time tc {t_tc*1e3:8.1f} us ({fl/t_tc/1e9:7.1f} TF/s)   simt {t_si*1e3:8.1f} us ({fl/t_si/1e9:6.1f} TF/s)", flush=True)
lib.gib_set_tensor_cores(1)
