"""GPU: capacity mode (device-side bond counts, no host synchronisation), int8 batches read directly by K0, and the
CUDA-graph training step -- each against the exact-size eager path of the same library."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(model="GGNN", B=96, seed=3, **kw):
    from graphinvent_b200 import synthetic as S
    from graphinvent_b200.gnn import mpnn
    from oracle import mpnn_oracle as O
    C = O.make_constants(model, **kw)
    sd = O.init_state_dict(C, seed=0)
    net = mpnn.create(C)
    net.load_state_dict(sd)
    net = net.cuda()
    n, e = S.random_graphs(B - 5, C.max_n_nodes, 5, 3, seed=seed, min_atoms=0)
    n2, e2 = S.corner_case_graphs(C.max_n_nodes, C.n_node_features)
    nodes = torch.from_numpy(np.concatenate([n2, n])).float().cuda()
    edges = torch.from_numpy(np.concatenate([e2, e])).float().cuda()
    apd = C.max_n_nodes * (C.len_f_add_per_node + C.len_f_conn_per_node) + 1
    target = torch.from_numpy(S.random_targets(B, apd, seed=seed + 1)).cuda()
    return C, net, nodes, edges, target


def _step_grads(net, nodes, edges, target):
    from graphinvent_b200 import functional as Fn
    net.zero_grad(set_to_none=True)
    out = net(nodes, edges)
    loss = Fn.kl_loss(out, target)
    loss.backward()
    return out.detach().clone(), float(loss.detach()), [p.grad.detach().clone() for p in net.parameters()]


@pytest.mark.parametrize("model", ["GGNN", "MNN", "AttGGNN"])
def test_capacity_mode_equals_exact_mode(model):
    C, net, nodes, edges, target = _setup(model)
    out0, loss0, g0 = _step_grads(net, nodes, edges, target)
    entries = net.last_stats["entries"]
    assert net.last_stats["capacity"] is None
    net.entry_capacity = int(entries * 1.3) + 64
    out1, loss1, g1 = _step_grads(net, nodes, edges, target)
    assert net.last_stats["capacity"] == net.entry_capacity
    # same tiles, same arithmetic: the logits agree to the last bit; weight gradients only differ by the split points
    # of the fixed-order reduction
    assert torch.equal(out0, out1)
    assert abs(loss0 - loss1) <= 1e-7
    for a, b in zip(g0, g1):
        assert (a - b).abs().max().item() <= 2e-6 * max(1e-3, a.abs().max().item())


def test_capacity_overflow_is_flagged_not_fatal():
    from graphinvent_b200 import functional as Fn
    C, net, nodes, edges, target = _setup("GGNN")
    _step_grads(net, nodes, edges, target)
    entries = net.last_stats["entries"]
    net.entry_capacity = max(8, entries // 3)
    graph = Fn.build_graph(net, edges)
    assert graph.overflowed()
    with torch.no_grad():
        out = net(nodes, edges, graph=graph)          # runs (truncated), must not fault
    torch.cuda.synchronize()
    assert out.shape[0] == nodes.shape[0]
    net.entry_capacity = entries + 1
    assert not Fn.build_graph(net, edges).overflowed()


@pytest.mark.parametrize("model", ["GGNN", "AttGGNN", "EMN"])
def test_int8_batches_equal_float_batches(model):
    """K0 and the first-layer kernels read the reference's on-disk int8 directly (DataProcesser.py:157-161)"""
    C, net, nodes, edges, target = _setup(model)
    out0, loss0, g0 = _step_grads(net, nodes, edges, target)
    out1, loss1, g1 = _step_grads(net, nodes.to(torch.int8), edges.to(torch.int8), target)
    assert torch.equal(out0, out1) and loss0 == loss1
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.int8])
def test_graphed_train_step_matches_eager_steps(in_dtype):
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200.graphed import TrainStep
    from graphinvent_b200.optim import FlatAdam
    C, net, nodes, edges, target = _setup("GGNN", B=128)
    net2 = copy.deepcopy(net)
    opt = FlatAdam(net.parameters(), lr=1e-4)          # (1e-3 makes this random-init model diverge: chaotic losses)
    opt2 = FlatAdam(net2.parameters(), lr=1e-4)
    # eager reference: the module API
    losses = []
    for _ in range(4):
        out = net(nodes, edges)
        loss = Fn.kl_loss(out, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    entries = net.last_stats["entries"]
    step = TrainStep(net2, opt2, batch_size=nodes.shape[0], entry_capacity=int(entries * 1.2) + 32, input_dtype=in_dtype)
    got = []
    for _ in range(4):
        got.append(float(step(nodes.to(in_dtype).cpu().pin_memory(), edges.to(in_dtype).cpu().pin_memory(), target)))
    assert step.check() & 4 == 0
    # the captured step differs from the eager one only in the split points of the weight-gradient reductions
    assert np.allclose(got, losses, rtol=0, atol=1e-5), (got, losses)
    for a, b in zip(net.parameters(), net2.parameters()):
        assert (a - b).abs().max().item() <= 1e-4
    # a batch that does not fit the capacity is reported
    small = TrainStep(net2, opt2, batch_size=nodes.shape[0], entry_capacity=max(8, entries // 4), input_dtype=in_dtype)
    small(nodes.to(in_dtype), edges.to(in_dtype), target)
    with pytest.raises(RuntimeError, match="entry_capacity"):
        small.check()
