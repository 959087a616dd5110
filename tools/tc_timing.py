"""GPU: where do the tcgen05 GEMM CTAs spend their cycles?  (gib_tc_timing; diagnosis build of tc_gemm_nt_kernel)

    python tools/tc_timing.py              # standalone shapes + one full C2 training step
    python tools/tc_timing.py --split2     # same with the SPLIT = 2 build (hi operand = raw tile)
    python tools/tc_timing.py --mask 128   # any gib_tc_debug mask, e.g. the CTA-pair kernel (its standalone GEMMs go
                                           # through gib_linear_fwd_tc_planes: the pair kernel needs weight planes)

Per role (one representative thread each) the kernel accumulates clock64 totals: time blocked on each mbarrier and
time doing its own work.  Printed per mode (NT = forward/dX launches, TN = weight-gradient launches) as a share of
the CTA's kernel time, mean over CTAs, plus cycles per k-block.  Reading it:
  TMA wait-for-free-stage high      -> consumers (split + MMA) are the bottleneck, loads are ahead
  splitter wait-for-TMA high        -> L2 / TMA latency-bound (more stages, multicast, larger boxes)
  MMA wait-for-split high + splitter work high -> the splitter is the bottleneck (pre-split activations, SPLIT = 2)
  MMA wait-for-accumulator high     -> the epilogue is the bottleneck
The numbers perturb the kernel a little (clock64 + a few registers); compare with the product timing printed beside it.
"""
import argparse
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from graphinvent_b200 import functional as Fn  # noqa: E402
from graphinvent_b200._lib import check, lib  # noqa: E402
from graphinvent_b200.gnn import mpnn  # noqa: E402
from graphinvent_b200.optim import FlatAdam  # noqa: E402

SLOTS = ["tma_wait_empty", "tma_total", "mma_wait_split", "mma_wait_acc", "mma_total", "spl_wait_raw", "spl_work",
         "spl_total", "epi_wait_acc", "epi_work", "epi_total", "kernel", "items", "kblocks", "launches", "-"]
CTAS = 160
dev = torch.device("cuda", 0)
P = lambda t: ctypes.c_void_p(t.data_ptr())


def report(buf, title):
    torch.cuda.synchronize()
    T = buf.view(2, CTAS, 16).double().cpu()
    for mode, name in ((0, "NT"), (1, "TN")):
        t = T[mode]
        live = t[:, 14] > 0
        if not live.any():
            continue
        t = t[live]
        k = t[:, 11]
        share = lambda i: 100.0 * (t[:, i] / k).mean().item()
        kb = t[:, 13].sum().item()
        print(f"[{title}] {name}: {int(live.sum())} CTAs, {int(t[:, 14].max())} launches, {int(t[:, 12].sum())} items, "
              f"{int(kb)} k-blocks, {k.sum().item() / max(kb, 1):.0f} CTA-cycles per k-block")
        print(f"    TMA   : wait-for-free-stage {share(0):5.1f} %  (loop {share(1):5.1f} % of kernel)")
        print(f"    split : wait-for-TMA        {share(5):5.1f} %  work {share(6):5.1f} %  -> {t[:, 6].sum().item() / max(kb, 1):.0f} cyc / k-block")
        print(f"    MMA   : wait-for-split      {share(2):5.1f} %  wait-for-accumulator {share(3):5.1f} %  issue+rest {share(4) - share(2) - share(3):5.1f} %")
        print(f"    epi   : wait-for-acc        {share(8):5.1f} %  work {share(9):5.1f} %  -> {t[:, 9].sum().item() / max(t[:, 12].sum().item(), 1):.0f} cyc / tile")
        print(f"    CTA kernel time: mean {k.mean().item():.0f}  min {k.min().item():.0f}  max {k.max().item():.0f} cycles")


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--split2", action="store_true")
    ap.add_argument("--mask", type=int, default=0)
    args = ap.parse_args()
    mask = args.mask | (64 if args.split2 else 0)
    lib.gib_tc_debug(mask)
    buf = torch.zeros(2 * CTAS * 16, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- standalone forward GEMMs (raw weights split in the kernel: gib_linear_fwd_tc) ------------------------
    for M, N, K in ((23808, 256, 256), (23808, 512, 512), (155648, 256, 256), (13312, 128, 128)):
        X = torch.randn(M, K, device=dev)
        W = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        Y = torch.empty(M, N, device=dev)
        if mask & 128:
            hi = ((W.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)                 # cvt.rna.tf32
            lo = (((W - hi).view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
            call = lambda: check(lib.gib_linear_fwd_tc_planes(P(X), K, P(hi), P(lo), K, P(b), P(Y), N, M, N, K, 1, st),
                                 "linear")
        else:
            call = lambda: check(lib.gib_linear_fwd_tc(P(X), K, P(W), K, P(b), P(Y), N, M, N, K, 1, st), "linear")
        lib.gib_tc_timing(None)
        ms = ev(call)
        ref = torch.nn.functional.selu(X.double() @ W.double().t() + b.double())
        err = (Y.double() - ref).abs().max().item()
        print(f"\nM={M} N={N} K={K}: product build {1e3 * ms:.1f} us = {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s, "
              f"max abs err vs fp64 {err:.2e}")
        buf.zero_()
        lib.gib_tc_timing(P(buf))
        ms_t = ev(call, K=5, warm=0)
        lib.gib_tc_timing(None)
        print(f"    timing build {1e3 * ms_t:.1f} us")
        report(buf, f"{M}x{N}x{K}")

    # ---- one full C2 training step (pre-split weights; NT and TN launches of the model) ------------------------
    C, nodes_h, edges_h, target_h, apd = bench.make_batch("C2", 1002)
    nodes, edges, target = nodes_h.to(dev), edges_h.to(dev), target_h.to(dev)
    torch.manual_seed(0)
    net = mpnn.create(C).to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-4)

    def step():
        loss = Fn.kl_loss(net(nodes, edges), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    lib.gib_tc_timing(None)
    ms = ev(step, K=10)
    print(f"\nC2 training step, product build: {ms:.3f} ms")
    buf.zero_()
    lib.gib_tc_timing(P(buf))
    ms_t = ev(step, K=3, warm=0)
    lib.gib_tc_timing(None)
    print(f"C2 training step, timing build: {ms_t:.3f} ms")
    report(buf, "C2 step x3")


if __name__ == "__main__":
    main()
