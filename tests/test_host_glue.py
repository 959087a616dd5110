"""Host logic of graphinvent_b200.functional (argument plumbing, caches, autograd wiring) on CPU tensors with the
library played by tests/hostshim.py -- the kernels themselves are covered by the `-m gpu` tests."""
import copy

import pytest
import torch

from tests import hostshim
from tests.conftest import MODELS, load_small


@pytest.fixture(params=MODELS)
def net_and_lib(request, monkeypatch):
    from graphinvent_b200.gnn import mpnn
    fx = load_small(request.param)
    net = mpnn.create(fx["C"])
    net.load_state_dict(fx["sd"])
    return net, hostshim.install_model_shims(monkeypatch, net), fx


def test_forward_backward_plumbing_and_flat_gradient_bucket(net_and_lib):
    net, lib, fx = net_and_lib
    seen = []
    net._grad_hook = lambda flat: seen.append(flat)
    out = net(fx["nodes"], fx["edges"])
    assert out.shape == (fx["nodes"].shape[0], fx["logits"].shape[1]) and out.requires_grad
    assert [c for c in lib.calls if c in ("gib_graph_count", "gib_graph_fill", "gib_model_pack", "gib_model_forward")] == \
        ["gib_graph_count", "gib_graph_fill", "gib_model_pack", "gib_model_forward"]
    out.sum().backward()
    assert lib.count("gib_model_backward") == 1 and len(seen) == 1
    ps = list(net.parameters())
    flat = seen[0]
    assert flat.numel() == sum(p.numel() for p in ps)
    off = 0
    for p in ps:                                   # every .grad is a view of the one bucket, in parameter order
        assert p.grad.shape == p.shape and p.grad.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()


def test_packed_weights_are_cached_until_a_parameter_changes(net_and_lib):
    from graphinvent_b200 import functional as Fn
    net, lib, fx = net_and_lib
    with torch.no_grad():
        net(fx["nodes"], fx["edges"])
        net(fx["nodes"][:3], fx["edges"][:3])                    # other batch size, same weights
        assert lib.count("gib_model_pack") == 1
        next(net.parameters()).add_(1.0)                         # in-place update bumps the version counter
        net(fx["nodes"], fx["edges"])
        assert lib.count("gib_model_pack") == 2
        Fn.invalidate_packed_weights()                           # writers behind autograd's back (flat Adam kernel)
        net(fx["nodes"], fx["edges"])
        assert lib.count("gib_model_pack") == 3
        twin = copy.deepcopy(net)
        twin(fx["nodes"], fx["edges"])                           # other tensors: own arena
        assert lib.count("gib_model_pack") == 4


def test_shared_graph_between_models_of_one_family(net_and_lib):
    from graphinvent_b200 import functional as Fn
    net, lib, fx = net_and_lib
    twin = copy.deepcopy(net)
    nodes, edges = fx["nodes"], fx["edges"]
    graph = Fn.build_graph(net, edges)
    k0 = lib.count("gib_graph_count")
    a = net(nodes, edges, graph=graph)
    b = twin(nodes, edges, graph=graph)
    assert lib.count("gib_graph_count") == k0 and lib.count("gib_model_forward") == 2
    (a.sum() + b.sum()).backward()
    assert lib.count("gib_model_backward") == 2
    with pytest.raises(ValueError, match="shared GraphBatch"):
        net(nodes, edges.clone(), graph=graph)                   # other tensor
    assert net._graph_in is None                                 # the hand-over slot is cleared even on errors
    with pytest.raises(ValueError, match="shared GraphBatch"):
        net(nodes[:4], edges[:4], graph=graph)                   # other batch size
    edges2 = edges.clone()
    g2 = Fn.build_graph(net, edges2)
    edges2[0, 0, 1, 0] = 1.0                                     # edited after K0 ran
    with pytest.raises(ValueError, match="shared GraphBatch"):
        net(nodes, edges2, graph=g2)
    with pytest.raises(ValueError):
        Fn.build_graph(net, edges.double())
    net(nodes, edges)                                            # and without a graph K0 runs as before
    assert lib.count("gib_graph_count") == k0 + 2


def test_no_grad_and_frozen_parameters_skip_autograd(net_and_lib):
    net, lib, fx = net_and_lib
    with torch.no_grad():
        assert not net(fx["nodes"], fx["edges"]).requires_grad
    for p in net.parameters():
        p.requires_grad_(False)
    assert not net(fx["nodes"], fx["edges"]).requires_grad


def test_shape_and_dropout_errors(net_and_lib):
    net, lib, fx = net_and_lib
    with pytest.raises(ValueError):
        net(fx["nodes"][0], fx["edges"])
    for m in net.modules():
        if hasattr(m, "dropout_p"):
            m.dropout_p = 0.1
    net.train()
    with pytest.raises(NotImplementedError):
        net(fx["nodes"], fx["edges"])
    net.eval()
    net(fx["nodes"], fx["edges"])                                # AlphaDropout is the identity in eval mode
