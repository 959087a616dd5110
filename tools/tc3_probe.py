"""GPU: where the second-generation tcgen05 GEMM spends its time -- the kernel with one pipeline role switched off at
a time (gib_tc_debug bits 8..15; results are wrong with most switches, timings only).  Launches are replayed from a
CUDA graph so that short kernels are not host-bound.
    python tools/tc3_probe.py [MxNxK ...]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
VARIANTS = [(0, "product"), (32768, "twelve N=128 MMAs per k-block"),
            (65536, "splitters with the software pipeline"), (1, "no global stores"), (2, "no epilogue after the drain"), (2 | 64, "no drain, no epilogue"),
            (4, "no split / STTM"), (8, "no W_lo tile and MMAs"), (16, "no MMAs"), (32, "no TMA loads"),
            (4 | 16 | 2 | 64, "TMA only"), (32 | 4 | 2 | 64, "MMA only"), (32 | 4 | 2 | 64 | 256, "MMA only, interleaved"),
            (256, "MMAs interleaved"), (512, "round-to-nearest activation split"),
            (32 | 4 | 2 | 64 | 2048, "MMA only, three accumulators"), (32 | 4 | 2 | 64 | 4096, "MMA only, N = 256 instructions"),
            (8192, "epilogue without the smem transpose"), (16384, "epilogue without the activation math"),
            (8192 | 16384, "epilogue: drain + stores only")]
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(155648, 256, 256), (23808, 256, 256),
                                                                         (23808, 256, 128), (13312, 512, 512)]


def graph_time(fn, n=20, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (n * reps)


for (M, N, K) in shapes:
    torch.manual_seed(0)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    check(lib.gib_split_planes(P(W), P(hi), P(lo), W.numel(), st()), "split")
    Y = torch.empty(M, N, device="cuda")
    G = torch.randn(M, N, device="cuda")
    dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    fl = 2.0 * M * N * K
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    print(f"== NT {M}x{N}x{K}: {tiles} tiles = {tiles / 148:.2f} waves, {K // 32} k-blocks per tile", flush=True)
    for mask, name in VARIANTS:
        lib.gib_tc_debug(mask << 8)
        try:
            t = graph_time(lambda: check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1,
                                                                      None, None, st()), "nt"))
            print(f"   {name:34s} {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TF/s", flush=True)
        except Exception as ex:
            print(f"   {name:34s} FAILED {ex}", flush=True)
            torch.cuda.synchronize()
    lib.gib_tc_debug(1)
    t = graph_time(lambda: check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, None,
                                                              None, st()), "nt1"))
    print(f"   {'first generation':34s} {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TF/s", flush=True)
    lib.gib_tc_debug(0)
    print(f"== TN (weight gradient + reduction) {M}x{N}x{K}", flush=True)
    for mask, name in [(0, "product"), (32768, "twelve N=128 MMAs per k-block"),
                       (65536, "splitters with the software pipeline"), (4, "no split / STTM"), (16, "no MMAs"), (32, "no TMA loads"), (2 | 64, "no drain, no epilogue"),
                       (512, "round-to-nearest activation split"), (1024, "no proxy fence after the X split")]:
        lib.gib_tc_debug(mask << 8)
        try:
            t = graph_time(lambda: check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), None,
                                                               None, st()), "dw"), n=10)
            print(f"   {name:34s} {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TF/s", flush=True)
        except Exception as ex:
            print(f"   {name:34s} FAILED {ex}", flush=True)
            torch.cuda.synchronize()
    lib.gib_tc_debug(0)
print("done")
