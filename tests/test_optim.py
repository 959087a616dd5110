"""
Flat-bucket Adam (graphinvent_b200.optim.FlatAdam / gib_adam_step) against torch.optim.Adam, the optimizer the
reference constructs at Workflow.py:191,221,245 and steps at Workflow.py:795-796.
Oracle here = torch.optim.Adam's single-tensor CPU implementation (the third-party arithmetic the reference calls).
"""
import copy

import pytest
import torch

from tests.conftest import load_small


def test_cpu_parameters_fail_loudly():
    from graphinvent_b200.optim import FlatAdam
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlatAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
    with pytest.raises(ValueError):
        FlatAdam([torch.nn.Parameter(torch.zeros(4))], lr=-1.0)


def test_flat_adam_host_logic_on_a_cpu_shim(monkeypatch):
    """bucket layout, run detection (missing gradients, lagging step counts, scattered vs bucket gradients), lr /
    betas pick-up, state_dict round trip and re-flattening after `.data` was re-assigned -- with the kernel played
    by a numpy restatement (tests/hostshim.py); reference = torch.optim.Adam(foreach=False) on CPU"""
    from tests import hostshim
    from graphinvent_b200.optim import FlatAdam
    launches = hostshim.install_optim_shims(monkeypatch)
    gen = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (3,), (33, 17), (1,), (250, 100), (625,)]
    init = _random_params(gen, shapes)
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    our_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=0.01, foreach=False)
    ours = FlatAdam(our_p, lr=1e-3, weight_decay=0.01)
    assert ours._in_place() and [p.data_ptr() for p in our_p] == [ours._flat.data_ptr() + 4 * o for o in ours._off]
    sched_r = torch.optim.lr_scheduler.OneCycleLR(ref, max_lr=1e-2, total_steps=12, pct_start=0.3)
    sched_o = torch.optim.lr_scheduler.OneCycleLR(ours, max_lr=1e-2, total_steps=12, pct_start=0.3)

    def close():
        return max((a.detach() - b.detach()).abs().max().item() for a, b in zip(ref_p, our_p))

    for step in range(5):                                        # scattered gradients, tensor 3 joins late
        for i, (a, b, g) in enumerate(zip(ref_p, our_p, _random_params(gen, shapes))):
            a.grad, b.grad = (None, None) if (i == 3 and step < 2) else (g.clone(), g.clone())
        del launches[:]
        ref.step(), ours.step(), sched_r.step(), sched_o.step()
        assert ours.grad_copies_last_step == (5 if step < 2 else 6)
        assert len(launches) == ours.launches_last_step == (2 if step < 2 else 3)
        assert ref.param_groups[0]["lr"] == ours.param_groups[0]["lr"]
        assert close() <= 1e-6, step
    total = sum(t.numel() for t in init)
    bucket = torch.randn(total, generator=gen)                   # the fused backward's layout: one flat bucket
    o = 0
    for a, b in zip(ref_p, our_p):
        a.grad = bucket[o:o + a.numel()].view(a.shape).clone()
        b.grad = bucket[o:o + a.numel()].view(a.shape)
        o += a.numel()
    del launches[:]
    ref.step(), ours.step()
    assert ours.grad_copies_last_step == 0 and launches == [7 * 5 + 3 + 33 * 17, 1, 250 * 100 + 625]
    assert close() <= 1e-6
    sd = copy.deepcopy(ours.state_dict())                        # same layout as torch.optim.Adam's
    rsd = ref.state_dict()
    assert sd["param_groups"][0]["params"] == rsd["param_groups"][0]["params"]
    for k in rsd["state"]:
        assert float(rsd["state"][k]["step"]) == float(sd["state"][k]["step"])
        assert (rsd["state"][k]["exp_avg_sq"] - sd["state"][k]["exp_avg_sq"]).abs().max().item() <= 1e-7
    for p in our_p:                                              # e.g. model.to(...) re-assigned the storage
        p.data = p.data.clone()
    assert not ours._in_place()
    for a, b, g in zip(ref_p, our_p, _random_params(gen, shapes)):
        a.grad, b.grad = g.clone(), g.clone()
    ref.step(), ours.step()
    assert ours._in_place() and close() <= 1e-6                   # re-flattened, moments and step counts kept
    again = FlatAdam(our_p, lr=1e-3, weight_decay=0.01)
    again.load_state_dict(sd)
    assert again._steps == [int(sd["state"][k]["step"]) for k in range(len(our_p))]


def _random_params(gen, shapes):
    return [torch.randn(*s, generator=gen) for s in shapes]


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_flat_adam_matches_torch_adam_on_raw_tensors(wd):
    """odd sizes (unaligned tensor starts inside the bucket), changing lr / betas between steps (OneCycleLR does
    both), a tensor without gradient in the middle (skipped, as torch does) and scattered gradient tensors"""
    from graphinvent_b200.optim import FlatAdam
    gen = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (3,), (33, 17), (1,), (250, 100), (625,)]
    init = _random_params(gen, shapes)
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=wd, foreach=False)
    ours = FlatAdam(our_p, lr=1e-3, weight_decay=wd)
    for step in range(6):
        lr, b1 = 1e-3 * (1 + step), 0.95 - 0.02 * step
        for o in (ref, ours):
            o.param_groups[0]["lr"] = lr
            o.param_groups[0]["betas"] = (b1, 0.999)
        grads = _random_params(gen, shapes)
        for i, (a, b, g) in enumerate(zip(ref_p, our_p, grads)):
            if i == 3 and step < 2:
                a.grad = b.grad = None
                continue
            a.grad = g.clone()
            b.grad = g.clone().cuda()            # separate allocations -> the gather-copy path
        ref.step()
        ours.step()
        assert ours.grad_copies_last_step > 0
        for a, b in zip(ref_p, our_p):
            assert (a.detach() - b.detach().cpu()).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item()), step
    # contiguous bucket path: gradients handed over as views of one flat buffer -> one launch per step-count run
    flat = torch.randn(sum(t.numel() for t in init), generator=gen)
    dflat = flat.cuda()
    o = 0
    for a, b in zip(ref_p, our_p):
        n = a.numel()
        a.grad = flat[o:o + n].view(a.shape).clone()
        b.grad = dflat[o:o + n].view(b.shape)
        o += n
    ref.step()
    ours.step()
    assert ours.grad_copies_last_step == 0 and ours.launches_last_step == 3    # tensor 3 is two steps behind
    for a, b in zip(ref_p, our_p):
        assert (a.detach() - b.detach().cpu()).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())
    # optimizer state_dict round trip keeps moments and step counts
    sd = copy.deepcopy(ours.state_dict())
    again = FlatAdam(our_p, lr=1e-3, weight_decay=wd)
    again.load_state_dict(sd)
    assert again._steps == ours._steps
    assert torch.equal(again._m, ours._m) and torch.equal(again._v, ours._v)
    rsd = ref.state_dict()
    for k in rsd["state"]:
        assert float(rsd["state"][k]["step"]) == float(sd["state"][k]["step"])
        assert (rsd["state"][k]["exp_avg"] - sd["state"][k]["exp_avg"].cpu()).abs().max().item() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["GGNN", "EMN"])
def test_training_with_flat_adam_follows_torch_adam(model):
    """three Workflow.train_epoch steps (Workflow.py:785-796) with OneCycleLR (Workflow.py:196-206): the drop-in
    optimizer and torch.optim.Adam give the same weights; one Adam launch per step, no gradient copies; the packed
    weight arena is refreshed although the kernel writes behind autograd's version counters"""
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200.gnn import mpnn
    from graphinvent_b200.optim import FlatAdam
    fx = load_small(model)
    nodes, edges, tgt = fx["nodes"].cuda(), fx["edges"].cuda(), fx["target"].cuda()

    def build():
        net = mpnn.create(fx["C"])
        net.load_state_dict(fx["sd"])
        return net.cuda()

    a, b = build(), build()
    oa = torch.optim.Adam(a.parameters(), lr=1e-4)
    ob = FlatAdam(b.parameters(), lr=1e-4)
    kw = dict(max_lr=1e-3, total_steps=4, pct_start=0.3)
    sa = torch.optim.lr_scheduler.OneCycleLR(oa, **kw)
    sb = torch.optim.lr_scheduler.OneCycleLR(ob, **kw)
    assert set(b.state_dict()) == set(fx["sd"])                       # still the reference's names
    first = None
    for step in range(3):
        outs = []
        for net, opt, sch in ((a, oa, sa), (b, ob, sb)):
            net.zero_grad()
            out = net(nodes, edges)
            Fn.kl_loss(out, tgt).backward()
            opt.step()
            sch.step()
            outs.append(out.detach())
        assert ob.launches_last_step == 1 and ob.grad_copies_last_step == 0
        assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"]
        assert oa.param_groups[0]["betas"] == ob.param_groups[0]["betas"]
        assert (outs[0] - outs[1]).abs().max().item() <= 1e-4, step
        if first is None:
            first = outs[1]
        else:
            assert not torch.equal(first, outs[1])                    # the weights really moved
    # weights: compare the accumulated UPDATE per tensor.  Adam turns a round-off-level gradient into a full +-lr
    # step, and from step 2 on the two runs see weights that differ in the last bit, so single elements whose true
    # gradient is ~0 may move differently; the norm of the update may not.
    for (k, pa), pb in zip(a.state_dict().items(), b.state_dict().values()):
        p0 = fx["sd"][k].cuda()
        da, db = pa - p0, pb - p0
        assert (da - db).norm().item() <= 5e-2 * da.norm().item() + 1e-12, k
        assert (da - db).abs().max().item() <= 2.5e-3, k                # never more than the summed learning rates
    twin = copy.deepcopy(b)                                           # Workflow.py:187-188 after an optimizer exists
    with torch.no_grad():
        assert torch.equal(twin(nodes, edges), b(nodes, edges))
    # a checkpoint written by the reference loads into the re-pointed parameters in place
    b.load_state_dict(fx["sd"])
    assert ob._in_place()
    with torch.no_grad():
        assert (b(nodes, edges).cpu() - fx["logits"]).abs().max().item() <= 1e-4
