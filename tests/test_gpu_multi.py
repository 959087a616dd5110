"""GPU (>= 2 devices): data-parallel gradients over NCCL equal the single-GPU gradients of the whole batch
(SURVEY.md §8e correctness check).  Skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["GIB_ROOT"])
from graphinvent_b200 import functional as Fn, parallel, synthetic as S
from graphinvent_b200.config import make_constants, apd_length
from graphinvent_b200.gnn import mpnn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
C = make_constants("GGNN")
torch.manual_seed(0)
net = mpnn.create(C).cuda()
parallel.broadcast_parameters(net)
n, e = S.random_graphs(256, 13, 5, 3, seed=9, min_atoms=1)
t = S.random_targets(256, apd_length(C), seed=9)
nodes, edges, tgt = (torch.from_numpy(a).float().cuda() for a in (n, e, t))
def grads(model, sl):
    model.zero_grad()
    Fn.kl_loss(model(nodes[sl], edges[sl]), tgt[sl]).backward()
    return [p.grad.clone() for p in model.parameters()]
full = grads(net, slice(0, 256))                       # single-GPU reference on the whole batch
hook = parallel.GradAllReduce(net)
lo, hi = parallel.shard_bounds(256, rank, world)
hook.set_shard(hi - lo, 256)
dp = grads(net, slice(lo, hi))                         # sharded + one all-reduce
worst = max(((a - b).norm() / b.norm().clamp_min(1e-12)).item() for a, b in zip(dp, full))
if rank == 0:
    print(json.dumps({"worst_rel_l2": worst, "allreduce_calls": hook.calls, "bytes": hook.bytes}))
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_gradients_match_single_gpu(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GIB_ROOT=ROOT)
    world = min(torch.cuda.device_count(), 4)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    # same function, different batch split: fp32 rounding + SELU-kink flips only (norm-level tolerance)
    assert res["worst_rel_l2"] <= 2e-3, res
    assert res["allreduce_calls"] == 1


STEP_WORKER = r'''
import copy, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["GIB_ROOT"])
from graphinvent_b200 import parallel, synthetic as S
from graphinvent_b200.config import make_constants, apd_length
from graphinvent_b200.gnn import mpnn
from graphinvent_b200.graphed import TrainStep
from graphinvent_b200.optim import FlatAdam
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
C = make_constants("GGNN")
torch.manual_seed(0)
net = mpnn.create(C).cuda()
parallel.broadcast_parameters(net)
ref = copy.deepcopy(net)
G = 256
n, e = S.random_graphs(G, 13, 5, 3, seed=9, min_atoms=1)
t = S.random_targets(G, apd_length(C), seed=9)
nodes, edges, tgt = torch.from_numpy(n).cuda(), torch.from_numpy(e).cuda(), torch.from_numpy(t).cuda()
cap = int((edges != 0).sum()) + 64
lo, hi = parallel.shard_bounds(G, rank, world)
step = TrainStep(net, FlatAdam(net.parameters(), lr=1e-4), batch_size=hi - lo, entry_capacity=cap,
                 input_dtype=torch.int8, global_batch=G)
losses = [float(step(nodes[lo:hi], edges[lo:hi], tgt[lo:hi])) for _ in range(3)]      # this rank's share of the loss
step.check()
tot = torch.tensor(losses, device="cuda", dtype=torch.float64)
dist.all_reduce(tot)
out = None
if rank == 0:
    one = TrainStep(ref, FlatAdam(ref.parameters(), lr=1e-4), batch_size=G, entry_capacity=cap, input_dtype=torch.int8,
                    global_batch=G, group=False)
    ref_losses = [float(one(nodes, edges, tgt)) for _ in range(3)]
    worst = max((a - b).abs().max().item() for a, b in zip(net.parameters(), ref.parameters()))
    out = {"param_max_abs_diff": worst, "dp_losses": tot.tolist(), "single_losses": ref_losses}
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_captured_data_parallel_steps_match_single_gpu(tmp_path):
    """graphed.TrainStep with a process group (two captured graphs, readout all-reduce overlapped) over 2 ranks ==
    the same three optimizer steps of the whole batch on one GPU"""
    script = tmp_path / "step_worker.py"
    script.write_text(STEP_WORKER)
    env = dict(os.environ, GIB_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["param_max_abs_diff"] <= 1e-4, res     # Adam normalises the step: a gradient element near zero may move by ~lr
    for a, b in zip(res["dp_losses"], res["single_losses"]):
        assert abs(a - b) <= 2e-5, res
