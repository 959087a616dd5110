#!/usr/bin/env python
"""
bench.py -- molecular-graphs/sec of one GGNN training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2|C4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch of synthetic molecules: forward ->
KL loss (Workflow.py:833-860) -> backward -> (N>1: one gradient all-reduce) -> Adam update.
Workload at N=1 = BASELINE.json configs[1] ("C2"): GGNN hidden=message=128, 4 message passes,
batch 1024 synthetic 13-atom molecules (gdb13 chemistry: F=8, Ef=3, APD=625), random-init weights.
N>1: every rank processes its own 1024-molecule batch (weak scaling), gradients all-reduced.

One JSON line on stdout (rank 0).  `value` = whole-job graphs/s with inputs resident in HBM;
`e2e` = same step driven from pinned HOST buffers (H2D of nodes/edges/targets + D2H of the loss
inside the timed region); `roofline` = dominant kernel class (the dense GEMMs) from live CUDA-event
timings inside the timed region; `roofline_scatter` = the scatter-aggregate kernel (K2) timed
alone at the C4 shape where its working set exceeds L2; `cpu_baseline` = the oracle port of the
reference's CPU path on this host.  `--impl reference` times that CPU path as its own arm.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (constants overrides, per-GPU batch, atoms, atom types, charges, description)
    "C2": (dict(hidden_node_features=128, message_size=128, message_passes=4), 1024, 13, 5, 3,
           "GGNN hidden=128, 4 MP steps, batch=1024 synthetic 13-node graphs"),
    "C4": (dict(max_n_nodes=38, n_node_features=12, len_f_add_per_node=81), 4096, 38, 9, 3,
           "GGNN defaults ZINC-scale synthetic (max_n_nodes=38), batch=4096"),
    "C1": (dict(), 100, 13, 5, 3, "GGNN defaults gdb13 dims, batch=100 (reference plumbing size)"),
    # profiling configurations of the other model families (not bench lines of BASELINE.json's metric)
    "C3": (dict(model="AttGGNN", hidden_node_features=256, message_size=256, message_passes=6, max_n_nodes=40,
                n_node_features=12, len_f_add_per_node=81), 2048, 40, 9, 3,
           "AttentionGGNN hidden=256, 6 MP steps, batch=2048 synthetic 40-node graphs"),
    "C5T": (dict(model="EMN"), 1000, 13, 5, 3, "EMN defaults gdb13 dims, batch=1000 (training step of the C5 model)"),
}
UNIT = "graphs/s"
CPU_MICRO_BATCH = 256     # the reference's O(V*E) prologue cannot run the large configurations whole


def config_model(cfg):
    return CONFIGS[cfg][0].get("model", "GGNN")


def metric_name(cfg):
    return f"molecular-graphs/sec (train fwd+bwd) {config_model(cfg)}"


METRIC = metric_name("C2")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


def make_batch(cfg, seed):
    from graphinvent_b200 import synthetic as S
    from graphinvent_b200.config import make_constants
    kw, B, n_atoms, n_types, n_charges, _ = CONFIGS[cfg]
    kw = {k: v for k, v in kw.items() if k != "model"}
    C = make_constants(config_model(cfg), **kw)
    nodes, edges = S.random_graphs(B, C.max_n_nodes, n_types, n_charges, seed=seed)
    apd = C.max_n_nodes * (C.len_f_add_per_node + C.len_f_conn_per_node) + 1
    target = S.random_targets(B, apd, seed=seed)
    return C, torch.from_numpy(nodes).float(), torch.from_numpy(edges).float(), torch.from_numpy(target), apd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout=3.0):
        """block until nvidia-smi delivered its first sample (it needs 0.1-1 s to start streaming)"""
        t_end = time.monotonic() + timeout
        while self.proc is not None and not self.rows and time.monotonic() < t_end:
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self, t0=None, t1=None):
        """t0, t1 (time.monotonic): the timed region.  nvidia-smi needs a few 100 ms to start streaming, so it is
        started before the warm-up steps (same load); samples inside [t0, t1] are used when there are any, else every
        sample taken under load since the start (and `window` says so)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows, window = [r for _, r in self.rows], "warm-up + timed region"
        if t0 is not None and t1 is not None:
            inside = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15]
            if inside:
                rows, window = inside, "timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


# ------------------------------------------------------------------------------------------
# CPU path of the reference (oracle port), used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------
def _pick_cpu_threads(O, C, nodes, edges, target):
    """The reference sets no thread count (PyTorch default = all cores); on a many-core host that default
    oversubscribes its small ATen ops badly, so give the CPU arm its best: time one step of a 128-molecule
    slice at a few thread counts and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    sd = O.init_state_dict(C, seed=0)
    n, e, t = nodes[:128], edges[:128], target[:128]
    best = (float("inf"), ncpu)
    for c in cands:
        torch.set_num_threads(c)
        O.train_step_grads(sd, C, n, e, t)                      # warm-up at this thread count
        t0 = time.perf_counter()
        O.train_step_grads(sd, C, n, e, t)
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, c)
    return best[1]


def cpu_train_steps(cfg, steps, warmup, budget_s=None, seed=1002):
    from oracle import mpnn_oracle as O
    C, nodes, edges, target, _ = make_batch(cfg, seed)
    if cfg not in ("C1", "C2"):      # micro-batch: the dense [V, E] prologue of the reference is quadratic in the batch
        nodes, edges, target = nodes[:CPU_MICRO_BATCH], edges[:CPU_MICRO_BATCH], target[:CPU_MICRO_BATCH]
    torch.set_num_threads(_pick_cpu_threads(O, C, nodes, edges, target))
    sd = O.init_state_dict(C, seed=0)
    params = [v.clone().requires_grad_(True) for v in sd.values()]
    leaves = dict(zip(sd.keys(), params))
    opt = torch.optim.Adam(params, lr=1e-4)
    B = nodes.shape[0]
    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.forward(leaves, C, nodes, edges)
        loss = O.kl_loss(out, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if budget_s is not None and it >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    return B, times, float(loss.detach())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None         # under torchrun only rank 0 measures the CPU path
    B, times, loss = cpu_train_steps(args.config, args.steps, args.warmup)
    total = sum(times)
    value = B * len(times) / total
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": metric_name(args.config), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][5], "name": args.config, "step": "fwd+kl_loss+bwd+adam",
                       "impl": "oracle port of the reference CPU path (oracle/mpnn_oracle.py; the reference is pure "
                               "PyTorch and is not mounted on this box)", "torch": torch.__version__},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(times)} full steps of batch {B}", "os_cpu_count": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "final_loss": loss}
    return line


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def scatter_roofline(pk, iters=30):
    """K2 alone at the C4 single-GPU shape (V=155648 slots, E=352256 entries, msg=100 -> ld 112):
    working set 204 MB > 126 MB L2, so consecutive launches cannot hit in L2."""
    from graphinvent_b200._lib import check, lib
    S, E, width = 155648, 352256, 100
    ld = (width + 15) // 16 * 16
    g = torch.Generator(device="cpu").manual_seed(0)
    dst = torch.randint(0, S, (E,), generator=g).sort().values
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    ptr[1:] = torch.bincount(dst, minlength=S).cumsum(0).int()
    ent = torch.arange(E, dtype=torch.int32)              # messages stored dst-sorted (as K0 + the type groups give)
    msg = torch.randn(E, ld, device="cuda"); w = torch.ones(E, device="cuda")
    out = torch.empty(S, ld, device="cuda")
    ptr, ent = ptr.cuda(), ent.cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for _ in range(5):
        check(lib.gib_scatter_sum(P(out), P(msg), ld, P(ptr), P(ent), P(w), S, st), "scatter")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        check(lib.gib_scatter_sum(P(out), P(msg), ld, P(ptr), P(ent), P(w), S, st), "scatter")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    nbytes = E * ld * 4 + S * ld * 4 + (S + 1) * 4 + E * 8
    gbs = nbytes / ms / 1e6
    return {"kernel": "scatter_sum_kernel (K2)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
            "frac": gbs / pk["hbm"], "traffic": 208559360,
            "traffic_source": "profiles/r01_ncu_full_summary.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, one launch)", "peak_source": pk["source"] + " (copy, burst)",
            "shape": {"slots": S, "entries": E, "ld": ld}, "ms_per_launch": ms, "bytes_per_launch": nbytes,
            "l2": "working set 204 MB > 126 MB L2, no flush needed"}


def run_b200_arm(args):
    import torch.distributed as dist
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200 import parallel
    from graphinvent_b200._lib import check, lib
    from graphinvent_b200.gnn import mpnn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    C, nodes_h, edges_h, target_h, apd = make_batch(args.config, 1002 + rank)
    B = nodes_h.shape[0]
    torch.manual_seed(0)                      # identical random-init replicas on every rank
    net = mpnn.create(C).to(dev)
    hook = parallel.GradAllReduce(net) if world > 1 else None
    if args.optimizer == "flat":              # one gib_adam_step launch over the flat parameter / gradient buckets
        from graphinvent_b200.optim import FlatAdam
        opt = FlatAdam(net.parameters(), lr=1e-4)
    else:
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    nodes, edges, target = nodes_h.to(dev), edges_h.to(dev), target_h.to(dev)
    pin = [t.pin_memory() for t in (nodes_h, edges_h, target_h)]
    h2d = sum(t.numel() * t.element_size() for t in pin)

    def step(n, e, t):
        out = net(n, e)
        loss = Fn.kl_loss(out, t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()               # before the warm-up: nvidia-smi takes a while to start streaming
        sampler.wait_first()
    for _ in range(max(args.warmup, 3)):
        loss = step(nodes, edges, target)
    barrier()

    # ---- timed region 1: device-resident inputs (no instrumentation) ---------------------
    launches0 = lib.gib_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.monotonic()
    ev0.record()
    for _ in range(args.steps):
        loss = step(nodes, edges, target)
    ev1.record()
    barrier()
    t_region1 = time.monotonic()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = lib.gib_launch_count() - launches0
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    final_loss = float(loss.detach())

    # ---- timed region 1b: the same K steps with a CUDA-event pair around every GEMM / scatter launch (the live
    #      per-kernel-class durations behind `roofline`; ~600 event records per step cost a few % -> kept out of `value`)
    lib.gib_profile_enable(1)
    barrier()
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for _ in range(args.steps):
        loss = step(nodes, edges, target)
    pv1.record()
    barrier()
    ms_instr = max_over_ranks(pv0.elapsed_time(pv1))
    pms = (ctypes.c_double * 3)(); pwork = (ctypes.c_double * 3)(); pcnt = (ctypes.c_longlong * 3)()
    check(lib.gib_profile_collect(pms, pwork, pcnt), "profile_collect")
    lib.gib_profile_enable(0)

    # ---- timed region 2: end to end from pinned host buffers -------------------------------
    # Every step: H2D of that step's inputs from pinned host memory (non-blocking, on the step's stream, as the
    # reference does at Workflow.py:781-782) and D2H of the step's loss into pinned memory; the host reads the value one
    # step later so that it never stalls the launch queue (tools/e2e_probe.py: a blocking .item() per step costs 0.4 ms,
    # the 5 MB of H2D 0.06 ms).  All copies and the final synchronisation are inside the timed region.
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_losses = []
    for i in range(args.steps):
        n, e, t = (x.to(dev, non_blocking=True) for x in pin)
        loss_i = step(n, e, t)
        loss_host[i & 1].copy_(loss_i.detach(), non_blocking=True)      # D2H of this step's result
        loss_events[i & 1].record()
        if i > 0:                                                        # previous step's loss, read on the host
            loss_events[(i - 1) & 1].synchronize()
            host_losses.append(float(loss_host[(i - 1) & 1]))
    loss_events[(args.steps - 1) & 1].synchronize()
    host_losses.append(float(loss_host[(args.steps - 1) & 1]))
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    assert len(host_losses) == args.steps and all(np.isfinite(host_losses))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    pk = peaks()
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    # dominant kernel class = forward/dX GEMMs (class 0) + dW GEMMs (class 1): report the larger one
    cls = 0 if pms[0] >= pms[1] else 1
    gemm_tflops = pwork[cls] / pms[cls] / 1e9 if pms[cls] > 0 else 0.0
    tensor_peak = pk["bf16_sustained"] / 2.0 / 3.0     # TF32 dense = bf16/2; fp32-accurate 3xTF32 = /3
    roofline = {"kernel": ["tc_gemm_nt_kernel NT mode (tcgen05 3xTF32; forward + dX GEMMs)",
                           "tc_gemm_nt_kernel TN mode + split reductions (weight-gradient GEMMs)"][cls],
                "bound": "tensor", "achieved": gemm_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": gemm_tflops / tensor_peak, "traffic": None,
                "peak_source": pk["source"] + " bf16 sustained / 2 (TF32 rate) / 3 (fp32-accurate 3xTF32 issue)",
                "launches_timed": int(pcnt[cls]), "ms_in_class": pms[cls],
                "share_of_step": pms[cls] / ms_instr, "ms_per_step_instrumented": ms_instr / args.steps,
                "other_classes": {"gemm_nt_ms": pms[0], "gemm_dw_ms": pms[1], "scatter_ms": pms[2],
                                  "gemm_nt_tflops": pwork[0] / pms[0] / 1e9 if pms[0] else 0,
                                  "gemm_dw_tflops": pwork[1] / pms[1] / 1e9 if pms[1] else 0}}
    line = {"metric": metric_name(args.config), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][5], "name": args.config, "per_gpu_batch": B,
                       "global_batch": B * world, "step": "fwd+kl_loss+bwd" + ("+allreduce" if world > 1 else "") + "+adam",
                       "optimizer": "graphinvent_b200.optim.FlatAdam (1 launch)" if args.optimizer == "flat" else "torch.optim.Adam(fused=True)",
                       "parallelism": f"dp{world}", "weights": "random init (reference initialisers: xavier-uniform MLPs, PyTorch-default GRU), torch.manual_seed(0)",
                       "l2": "per-step working set (saved activations + packed weights, "
                             f"{net.last_stats.get('workspace_bytes', 0) / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                       "bond_entries_per_batch": net.last_stats.get("entries")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 + 64,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "public module API; pinned host batch -> non-blocking H2D -> step -> loss D2H to pinned memory, "
                           "read on the host one step later; +64 B/step = K0 graph header"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "final_loss": final_loss}
    try:
        line["roofline_scatter"] = scatter_roofline(pk)
    except Exception as ex:  # keep the headline line even if the side measurement fails
        line["roofline_scatter"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        Bc, times, _ = cpu_train_steps(args.config, steps=8, warmup=1, budget_s=20.0)
        v = Bc * len(times) / sum(times)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{len(times)} full steps of batch {Bc} after 1 warm-up (oracle port, "
                                          "fwd+kl_loss+bwd+adam)", "os_cpu_count": os.cpu_count()}
    if world > 1:
        dist.destroy_process_group()
    return line


class _StdoutToStderr:
    """NCCL (and anything else native) may print to fd 1 (e.g. "NCCL version ..."): the contract is ONE JSON line on
    stdout, so fd 1 is pointed at stderr while the benchmark runs and restored just before the line is printed."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat = graphinvent_b200.optim.FlatAdam (one launch); torch = torch.optim.Adam(fused=True)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 5
        args.warmup = args.warmup if args.warmup is not None else 1
        with _StdoutToStderr():
            line = run_reference_arm(args)
    else:
        args.steps = args.steps if args.steps is not None else 30
        args.warmup = args.warmup if args.warmup is not None else 5
        with _StdoutToStderr():
            line = run_b200_arm(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
