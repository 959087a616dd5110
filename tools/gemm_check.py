"""GPU: the tcgen05 3xTF32 GEMMs against fp64 -- second generation (activation operand through tensor memory,
gemm_tc3.cu) and first generation (gemm_tc.cu) side by side: accuracy, device-side row ranges, speed.

    python tools/gemm_check.py                # probes + accuracy + timing table
    python tools/gemm_check.py quick          # probes + accuracy only
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
quick = "quick" in sys.argv[1:]


def planes(W):
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    check(lib.gib_split_planes(P(W), P(hi), P(lo), W.numel(), st()), "split")
    return hi, lo


def nt(X, Whl, b, Y, M, N, K, act, gen, m_dev=None, base_dev=None):
    lib.gib_tc_debug({1: 1, 2: 0, 3: 32768 << 8}[gen])      # 3 = second generation with twelve N = 128 MMAs per k-block
    check(lib.gib_linear_fwd_tc_planes(P(X), X.shape[1], P(Whl[0]), P(Whl[1]), K, P(b), P(Y), Y.shape[1], M, N, K, act,
                                       P(m_dev), P(base_dev), st()), f"linear gen{gen}")
    lib.gib_tc_debug(0)


def dw(G, X, M, N, K, gen, m_dev=None, base_dev=None, sc=None):
    lib.gib_tc_debug({1: 1, 2: 0, 3: 32768 << 8}[gen])
    dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
    if sc is None:
        sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    check(lib.gib_linear_bwd_dw(P(G), G.shape[1], N, P(X), X.shape[1], K, M, P(dW), P(db), N, K, P(sc), P(m_dev),
                                P(base_dev), st()), f"dw gen{gen}")
    lib.gib_tc_debug(0)
    return dW, db, sc


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ok = True

# ---- 1. layout probe: W = identity -> Y must reproduce X element for element (a swizzle / lane / column mix-up shows
#         up as a permutation, which is printed) ----------------------------------------------------------------------
for (M, K) in [(128, 32), (128, 64), (256, 128), (300, 144)]:
    torch.manual_seed(1)
    X = torch.randn(M, K, device="cuda")
    W = torch.eye(K, device="cuda")
    Y = torch.full((M, K), float("nan"), device="cuda")
    try:
        nt(X, planes(W), None, Y, M, K, K, 0, gen=2)
        torch.cuda.synchronize()
    except Exception as ex:
        print(f"probe M={M} K={K}: LAUNCH FAILED {ex}", flush=True)
        ok = False
        break
    err = (Y - X).abs().max().item()
    print(f"probe identity M={M:4d} K={K:4d}: max |Y - X| = {err:.2e}", flush=True)
    if not err < 1e-5:
        ok = False
        # where did each element go?
        for m in (0, 1, 7, 8, 33, 127):
            row = []
            for k in range(min(K, 16)):
                hit = (Y - X[m, k]).abs() < 1e-6
                idx = hit.nonzero()
                row.append(tuple(idx[0].tolist()) if len(idx) else None)
            print(f"   X[{m}, 0:16] landed at {row}", flush=True)
        bad_rows = ((Y - X).abs().max(1).values > 1e-5).nonzero().flatten()[:16].tolist()
        bad_cols = ((Y - X).abs().max(0).values > 1e-5).nonzero().flatten()[:16].tolist()
        print(f"   bad rows {bad_rows} bad cols {bad_cols} nonfinite {int((~torch.isfinite(Y)).sum())}", flush=True)

# ---- 2. accuracy vs fp64, both generations --------------------------------------------------------------------------
shapes = [(128, 128, 32), (100, 48, 16), (1000, 256, 256), (23808, 256, 256), (23808, 128, 256), (23808, 256, 128),
          (13312, 512, 512), (13312, 384, 128), (4096, 48, 512), (5000, 608, 512), (13312, 256, 144), (1024, 500, 688),
          (155648, 256, 256)]
for (M, N, K) in shapes:
    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    Whl = planes(W)
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    line = f"NT M={M:6d} N={N:4d} K={K:4d}:"
    for gen in (2, 3, 1):
        for act in (0, 1):
            r = torch.selu(ref) if act else ref
            Y = torch.full((M, N), float("nan"), device="cuda")
            try:
                nt(X, Whl, b, Y, M, N, K, act, gen)
                torch.cuda.synchronize()
                e = (Y.double() - r).abs().max().item()
                bad = int((~torch.isfinite(Y)).sum())
            except Exception as ex:
                print(f"{line} gen{gen} act{act} FAILED {ex}", flush=True)
                e, bad = float("inf"), -1
            line += f" gen{gen}/act{act} err {e:.2e}" + (f" NONFINITE {bad}" if bad else "")
            if gen == 2 and not e < 2e-5:
                ok = False
    print(line, flush=True)
    if not quick:
        Y = torch.empty(M, N, device="cuda")
        t2 = timeit(lambda: nt(X, Whl, b, Y, M, N, K, 1, 2))
        t1 = timeit(lambda: nt(X, Whl, b, Y, M, N, K, 1, 1))
        fl = 2.0 * M * N * K
        print(f"        time gen2 {t2*1e3:8.1f} us ({fl/t2/1e9:7.1f} TF/s)   gen1 {t1*1e3:8.1f} us ({fl/t1/1e9:6.1f} TF/s)", flush=True)

# ---- 3. device-side row range (capacity mode) ------------------------------------------------------------------------
torch.manual_seed(5)
cap, N, K = 4096, 256, 256
X = torch.randn(cap, K, device="cuda")
X[3000:] = float("nan")                       # rows beyond the live range may hold anything
W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda")
Whl = planes(W)
for (base, m) in [(0, 1000), (128, 2500), (2944, 56), (0, 0)]:
    md = torch.tensor([m], dtype=torch.int32, device="cuda"); bd = torch.tensor([base], dtype=torch.int32, device="cuda")
    Y = torch.full((cap, N), 7.0, device="cuda")
    nt(X, Whl, b, Y, cap, N, K, 1, 2, md, bd)
    torch.cuda.synchronize()
    ref = torch.selu(torch.nn.functional.linear(X[base:base + m].double(), W.double(), b.double()))
    e = (Y[base:base + m].double() - ref).abs().max().item() if m else 0.0
    untouched = bool((Y[:base] == 7.0).all() and (Y[base + m:] == 7.0).all())
    print(f"NT dynamic rows base={base} m={m}: err {e:.2e}, rows outside untouched: {untouched}", flush=True)
    if not (e < 2e-5 and untouched):
        ok = False

# ---- 4. weight gradients ---------------------------------------------------------------------------------------------
shapes = [(2048, 128, 128), (4100, 256, 256), (23808, 256, 256), (23808, 256, 128), (13312, 512, 512), (13312, 384, 128),
          (13312, 48, 512), (13312, 256, 144), (20000, 608, 512), (155648, 256, 256)]
for (M, N, K) in shapes:
    torch.manual_seed(M + N)
    G = torch.randn(M, N, device="cuda"); X = torch.randn(M, K, device="cuda")
    ref = G.double().t() @ X.double(); refb = G.double().sum(0)
    line = f"TN M={M:6d} N={N:4d} K={K:4d}:"
    for gen in (2, 3, 1):
        try:
            dW, db, sc = dw(G, X, M, N, K, gen)
            torch.cuda.synchronize()
            e = (dW.double() - ref).abs().max().item(); eb = (db.double() - refb).abs().max().item()
        except Exception as ex:
            print(f"{line} gen{gen} FAILED {ex}", flush=True)
            e = eb = float("inf")
        scale = ref.abs().max().item()
        line += f" gen{gen} err dW {e:.2e} db {eb:.2e} (|dW|max {scale:.1f})"
        if gen == 2 and not (e < 3e-5 * max(1.0, scale) and eb < 1e-5 * max(1.0, refb.abs().max().item())):
            ok = False
        if not quick and e < float("inf"):
            t = timeit(lambda: dw(G, X, M, N, K, gen, sc=sc), 10)
            line += f" {t*1e3:7.1f} us ({2.0*M*N*K/t/1e9:5.1f} TF/s)"
    print(line, flush=True)

# TN with a device-side row range
cap, N, K = 8192, 256, 256
G = torch.randn(cap, N, device="cuda"); X = torch.randn(cap, K, device="cuda")
G[6000:] = float("nan"); X[6000:] = float("inf")
for (base, m) in [(0, 3000), (640, 5000), (5888, 100), (0, 0)]:
    md = torch.tensor([m], dtype=torch.int32, device="cuda"); bd = torch.tensor([base], dtype=torch.int32, device="cuda")
    dW, db, _ = dw(G, X, cap, N, K, 2, md, bd)
    torch.cuda.synchronize()
    ref = G[base:base + m].double().t() @ X[base:base + m].double(); refb = G[base:base + m].double().sum(0)
    e = (dW.double() - ref).abs().max().item(); eb = (db.double() - refb).abs().max().item()
    print(f"TN dynamic rows base={base} m={m}: err dW {e:.2e} db {eb:.2e}", flush=True)
    if not (e < 3e-5 * max(1.0, ref.abs().max().item() if m else 1.0) and eb < 1e-5 * max(1.0, refb.abs().max().item() if m else 1.0)):
        ok = False

print("GEMM_CHECK", "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
