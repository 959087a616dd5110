"""CPU: the oracle (oracle/mpnn_oracle.py) against the golden fixtures produced by the
unmodified reference, and against the live reference when /root/reference is mounted."""
import pytest
import torch

from oracle import mpnn_oracle as O
from tests import refimpl
from tests.conftest import MODELS, load_gdb13, load_small, pretrained_path

LOGIT_TOL = 1e-5      # fp32 re-association noise only (observed <= 1e-6)
GRAD_REL_TOL = 1e-5


@pytest.mark.parametrize("model", MODELS)
def test_oracle_matches_golden_logits_loss_grads(model):
    fx = load_small(model)
    loss, out, grads = O.train_step_grads(fx["sd"], fx["C"], fx["nodes"], fx["edges"], fx["target"])
    assert (out - fx["logits"]).abs().max().item() <= LOGIT_TOL
    assert torch.equal(out.argmax(1), fx["logits"].argmax(1))
    assert abs(float(loss) - fx["loss"]) <= 1e-6
    for k, g in fx["grads"].items():
        scale = max(g.abs().max().item(), 1e-12)
        assert (grads[k] - g).abs().max().item() / scale <= GRAD_REL_TOL, k


@pytest.mark.parametrize("model", MODELS)
def test_param_schema_matches_reference_state_dict(model):
    fx = load_small(model)
    shapes = O.param_shapes(fx["C"])
    assert [k for k, _ in shapes] == list(fx["sd"].keys())
    assert all(tuple(fx["sd"][k].shape) == tuple(s) for k, s in shapes)
    sd = O.init_state_dict(fx["C"], seed=1)
    assert list(sd.keys()) == list(fx["sd"].keys())


def test_oracle_pretrained_gdb13_rows():
    path = pretrained_path()
    if path is None:
        pytest.skip("tests/golden/_local/pretrained_model.pth absent (run tests/golden/make_golden.py)")
    fx = load_gdb13()
    sd = torch.load(path, map_location="cpu", weights_only=False)
    C = O.make_constants("GGNN")
    out = O.forward(sd, C, fx["nodes"], fx["edges"])
    assert (out - fx["logits"]).abs().max().item() <= 2e-5
    assert torch.equal(out.argmax(1), fx["logits"].argmax(1))
    loss = O.kl_loss(out, fx["apds"])
    assert abs(float(loss) - fx["loss"]) <= 1e-5


def test_kl_loss_matches_definition():
    torch.manual_seed(0)
    out = torch.randn(7, 33)
    t = torch.rand(7, 33)
    t[0, :5] = 0
    th = t / t.sum(1, keepdim=True)
    want = (torch.xlogy(th, th) - th * torch.log_softmax(out, 1)).sum() / 7
    assert abs(float(O.kl_loss(out, t)) - float(want)) < 1e-6


@pytest.mark.skipif(not refimpl.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("model", MODELS)
def test_oracle_matches_live_reference(model):
    from graphinvent_b200 import synthetic as S
    torch.manual_seed(5)
    fx = load_small(model)
    C = fx["C"]
    net = refimpl.build(C)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    n, e = S.random_graphs(24, C.max_n_nodes, 4, 2, seed=77, min_atoms=0)
    nodes, edges = torch.from_numpy(n).float(), torch.from_numpy(e).float()
    with torch.no_grad():
        ref = net(nodes, edges)
        out = O.forward(sd, C, nodes, edges)
    assert (out - ref).abs().max().item() <= LOGIT_TOL
