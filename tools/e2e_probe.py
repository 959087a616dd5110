"""GPU: where does the end-to-end step lose time vs the device-resident step? (bench.py e2e diagnosis)"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from graphinvent_b200 import functional as Fn
from graphinvent_b200.gnn import mpnn
C, nodes_h, edges_h, target_h, apd = bench.make_batch("C2", 1002)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = mpnn.create(C).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
nodes, edges, target = nodes_h.to(dev), edges_h.to(dev), target_h.to(dev)
pin = [t.pin_memory() for t in (nodes_h, edges_h, target_h)]
def step(n, e, t):
    out = net(n, e); loss = Fn.kl_loss(out, t); opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
def timeit(name, fn, K=30):
    for _ in range(3): fn(0)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(K): fn(i)
    b.record(); torch.cuda.synchronize()
    print(f"{name:45s} {a.elapsed_time(b)/K:7.3f} ms/step (wall {1e3*(time.perf_counter()-t0)/K:7.3f})", flush=True)
timeit("resident", lambda i: step(nodes, edges, target))
timeit("resident + item()", lambda i: step(nodes, edges, target).item())
timeit("h2d main stream", lambda i: step(*[t.to(dev, non_blocking=True) for t in pin]))
timeit("h2d main stream + item()", lambda i: step(*[t.to(dev, non_blocking=True) for t in pin]).item())
pre = [t.to(dev) for t in pin]
def copy_into(i):
    for d, s in zip(pre, pin): d.copy_(s, non_blocking=True)
    return step(*pre)
timeit("h2d copy_ into persistent device buffers", copy_into)
timeit("resident, fresh clones each step", lambda i: step(nodes.clone(), edges.clone(), target.clone()))
# host-side enqueue cost of one step (no sync except the K0 header read)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): step(nodes, edges, target)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue per step {1e3*(t1-t0)/10:.3f} ms, drained after {1e3*(t2-t1):.3f} ms")
