"""GPU: the CTA-pair tcgen05 GEMM (gemm_tc2.cu, gib_tc_debug bit 7) on single GEMMs, smallest shapes first.

    timeout 120 python tools/tc2_test.py              # forward (NT) shapes, then weight-gradient (TN) shapes
    timeout 120 python tools/tc2_test.py nt 256x128x32  # one shape

Each shape runs through the single-CTA kernel (mask 0) and the pair kernel (mask 128, and 192 = raw hi operand) via
gib_linear_fwd_tc_planes / gib_linear_bwd_dw and is compared with fp64.  The first shapes have one k-block and one
or two pair tiles, so a broken hand-off (trap after ~10 s) or a wrong descriptor shows up on the cheapest case.
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, on the 13 dropped mantissa bits"""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def nt(M, N, K):
    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    hi = tf32_rna(W)
    lo = tf32_rna(W - hi)
    ref = torch.nn.functional.selu(X.double() @ W.double().t() + b.double())
    for mask in (0, 128, 192):
        lib.gib_tc_debug(mask)
        Y = torch.full((M, N), float("nan"), device="cuda")
        call = lambda: check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, st()), "nt")
        call()
        torch.cuda.synchronize()
        err = (Y.double() - ref).abs().max().item()
        ms = timed(call)
        print(f"NT {M}x{N}x{K} mask {mask:3d}: max abs err {err:.2e}  {1e3 * ms:8.1f} us  {2.0 * M * N * K / ms / 1e9:6.1f} TFLOP/s",
              flush=True)
    lib.gib_tc_debug(0)


def tn(M, N, K):
    torch.manual_seed(M + N)
    G = torch.randn(M, N, device="cuda")
    X = torch.randn(M, K, device="cuda")
    ref = G.double().t() @ X.double()
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    for mask in (0, 128, 192):
        lib.gib_tc_debug(mask)
        dW = torch.zeros(N, K, device="cuda")
        db = torch.zeros(N, device="cuda")
        call = lambda: check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), st()), "tn")
        call()
        torch.cuda.synchronize()
        err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
        ms = timed(call)
        print(f"TN {M}x{N}x{K} mask {mask:3d}: max err / max|dW| {err:.2e}  {1e3 * ms:8.1f} us  {2.0 * M * N * K / ms / 1e9:6.1f} TFLOP/s",
              flush=True)
    lib.gib_tc_debug(0)


NT = [(256, 128, 32), (256, 128, 64), (128, 128, 256), (300, 64, 128), (1024, 256, 256), (23808, 256, 256),
      (23808, 512, 512), (13312, 128, 128), (155648, 256, 256)]
TN = [(2048, 256, 128), (2048, 128, 128), (4100, 256, 256), (23808, 256, 256), (13312, 512, 512), (155648, 256, 256)]
if len(sys.argv) > 2:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
    for s in shapes:
        (nt if sys.argv[1] == "nt" else tn)(*s)
else:
    for s in NT:
        nt(*s)
    for s in TN:
        tn(*s)
