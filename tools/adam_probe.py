"""GPU: cost of the flat Adam launch vs torch's fused Adam, and whether the flat parameter layout changes the step."""
import ctypes, sys, time, torch
sys.path.insert(0, ".")
import bench
from graphinvent_b200 import functional as Fn
from graphinvent_b200._lib import check, lib
from graphinvent_b200.gnn import mpnn
from graphinvent_b200.optim import FlatAdam

dev = torch.device("cuda", 0)


def ev_time(fn, K=30, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(K):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K, 1e3 * (time.perf_counter() - t0) / K


# ---- 1. the kernel alone ------------------------------------------------------------------------------------
n = 6_037_329
bufs = [torch.randn(n + 8, device=dev) for _ in range(4)]
bufs[3].abs_()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for sh in (0, 1):
    P = [ctypes.c_void_p(t.data_ptr() + 4 * sh) for t in bufs]
    ms, _ = ev_time(lambda i: check(lib.gib_adam_step(P[0], P[1], P[2], P[3], n, i + 1, 1e-4, 0.9, 0.999, 1e-8, 0.0,
                                                      1.0, st), "adam"), K=50)
    print(f"gib_adam_step n={n} shift={sh}: {1e3 * ms:8.1f} us  ({28 * n / ms / 1e6:7.1f} GB/s)", flush=True)

C, nodes_h, edges_h, target_h, apd = bench.make_batch("C2", 1002)
nodes, edges, target = nodes_h.to(dev), edges_h.to(dev), target_h.to(dev)


def make(kind):
    torch.manual_seed(0)
    net = mpnn.create(C).to(dev)
    if kind == "torch":
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    elif kind == "flat":
        opt = FlatAdam(net.parameters(), lr=1e-4)
    elif kind == "flat-layout+torch":
        FlatAdam(net.parameters(), lr=1e-4)              # only re-points the parameters into one bucket
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    elif kind == "none":
        opt = None
    elif kind == "flat-layout+none":
        FlatAdam(net.parameters(), lr=1e-4)
        opt = None
    return net, opt


for kind in ("torch", "flat", "flat-layout+torch", "none", "flat-layout+none", "torch", "flat"):
    net, opt = make(kind)

    def step(i):
        out = net(nodes, edges)
        loss = Fn.kl_loss(out, target)
        net.zero_grad(set_to_none=True)
        loss.backward()
        if opt is not None:
            opt.step()
        else:
            Fn.invalidate_packed_weights()               # keep the per-step weight packing in the picture
        return loss

    ms, wall = ev_time(step)
    # the optimizer call alone, gradients in place
    if opt is not None:
        oms, owall = ev_time(lambda i: opt.step(), K=50)
        print(f"{kind:20s} step {ms:7.3f} ms (wall {wall:7.3f})   optimizer.step alone {1e3 * oms:7.1f} us device, "
              f"{1e3 * owall:7.1f} us wall", flush=True)
    else:
        print(f"{kind:20s} step {ms:7.3f} ms (wall {wall:7.3f})", flush=True)
    del net, opt
    torch.cuda.empty_cache()
