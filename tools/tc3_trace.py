"""GPU: per-tile timeline of CTA 0 of the second-generation GEMM (gib_tc_trace): when does the MMA warp start / finish a
tile, how long does it wait for TMA tiles / the split operand, when does the epilogue see, drain and store the tile.
    python tools/tc3_trace.py [MxNxK] [tiles] [tn]      (tn: the weight-gradient kernel, one line per work item)"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from graphinvent_b200._lib import check, lib  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "155648x256x256").split("x"))
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 12
X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
hi, lo = torch.empty_like(W), torch.empty_like(W)
check(lib.gib_split_planes(P(W), P(hi), P(lo), W.numel(), st()), "split")
Y = torch.empty(M, N, device="cuda")
for mask, name in [(0, "product"), (1, "no global stores"), (2, "no epilogue after the drain")]:
    buf = torch.zeros(tiles, 16, dtype=torch.int64, device="cuda")
    lib.gib_tc_debug(mask << 8)
    for _ in range(2):
        check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, None, None, st()), "nt")
    lib.gib_tc_trace(P(buf), tiles)
    check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, None, None, st()), "nt")
    torch.cuda.synchronize()
    lib.gib_tc_trace(None, 0)
    lib.gib_tc_debug(0)
    t = buf.cpu()
    t0 = int(t[0, 0])
    print(f"== {name}: cycles relative to the first MMA start of CTA 0")
    print("tile  mma_start  mma_last  (wait tma, wait split)   epi_sees  drained   stored   | splitter wait / work")
    for i in range(tiles):
        r = t[i]
        print(f"{i:4d} {int(r[0]) - t0:10d} {int(r[1]) - t0:9d}  ({int(r[2]):6d}, {int(r[3]):6d})   {int(r[4]) - t0:9d} {int(r[5]) - t0:8d} {int(r[6]) - t0:8d}   | "
              f"{int(r[8]):6d} / {int(r[9]):6d}")

if "tn" in sys.argv[1:]:
    G = torch.randn(M, N, device="cuda")
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
    run = lambda: check(lib.gib_linear_bwd_dw(P(G), N, N, P(X), K, K, M, P(dW), P(db), N, K, P(sc), None, None, st()), "dw")
    for mask, name in [(0, "product"), (65536, "splitters with the software pipeline"), (4, "no split")]:
        buf = torch.zeros(tiles, 16, dtype=torch.int64, device="cuda")
        lib.gib_tc_debug(mask << 8)
        for _ in range(2):
            run()
        lib.gib_tc_trace(P(buf), tiles)
        run()
        torch.cuda.synchronize()
        lib.gib_tc_trace(None, 0)
        lib.gib_tc_debug(0)
        t = buf.cpu()
        t0 = int(t[0, 0])
        print(f"== TN {name}: cycles relative to the first MMA start of CTA 0")
        print("item  mma_start  mma_last  (wait tma, wait split)   epi_sees  drained   stored   | G splitter wait / work"
              " | X splitter wait / work")
        for i in range(tiles):
            r = t[i]
            if int(r[0]) == 0:
                break
            print(f"{i:4d} {int(r[0]) - t0:10d} {int(r[1]) - t0:9d}  ({int(r[2]):6d}, {int(r[3]):6d})   {int(r[4]) - t0:9d} "
                  f"{int(r[5]) - t0:8d} {int(r[6]) - t0:8d}   | {int(r[8]):6d} / {int(r[9]):6d} | {int(r[10]):6d} / {int(r[11]):6d}")
