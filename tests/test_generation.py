"""§8(f) rank 1: the device-side generation round against a trace of the unmodified reference generator.
The trace (tests/golden/generation_trace.npz, made by tests/golden/make_generation_trace.py) holds the draws the
reference sampled in every round and its final buffers; replaying the draws must reproduce the buffers bit-exactly."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN, pretrained_path


def _trace():
    return np.load(os.path.join(GOLDEN, "generation_trace.npz"))


def test_trace_fixture_is_self_consistent():
    z = _trace()
    B, n_gen, R = int(z["batch"]), int(z["n_generated"]), int(z["rounds"])
    assert z["actions"].shape == (R, B) and z["likelihoods"].shape == (R, B)
    assert B <= n_gen <= 2 * B and R < 26
    assert 0 <= z["actions"].min() and z["actions"].max() <= 624
    nn = z["generated_n_nodes"][:n_gen]
    atoms = (z["generated_nodes"][:n_gen].sum(-1) > 0).sum(-1)
    assert (atoms == nn).all()                                   # node rows agree with the atom counts
    assert (z["generated_edges"][:n_gen] == z["generated_edges"][:n_gen].transpose(0, 2, 1, 3)).all()
    assert z["properly_terminated"][:n_gen].sum() >= n_gen // 3  # trained checkpoint: most molecules terminate properly


@pytest.mark.gpu
def test_replay_reproduces_the_reference_generator_bit_exactly():
    from graphinvent_b200.config import make_constants
    from graphinvent_b200.generation import GraphGenerator
    z = _trace()
    B, n_gen = int(z["batch"]), int(z["n_generated"])
    C = make_constants("GGNN")
    gen = GraphGenerator(model=None, batch_size=B, constants=C, n_atom_types=5, n_formal_charge=3)
    replay = [(torch.from_numpy(a), torch.from_numpy(l)) for a, l in zip(z["actions"], z["likelihoods"])]
    got = gen.build_graphs(replay=replay)
    assert got == n_gen and gen.rounds == int(z["rounds"])
    assert torch.equal(gen.generated_nodes.cpu().to(torch.int8), torch.from_numpy(z["generated_nodes"]))
    assert torch.equal(gen.generated_edges.cpu().to(torch.int8), torch.from_numpy(z["generated_edges"]))
    assert torch.equal(gen.generated_n_nodes.cpu(), torch.from_numpy(z["generated_n_nodes"]))
    assert torch.equal(gen.generated_likelihoods.cpu(), torch.from_numpy(z["generated_likelihoods"]))
    assert torch.equal(gen.properly_terminated.cpu(), torch.from_numpy(z["properly_terminated"]))
    # the live batch state after the last round too (incl. the dummy graph's accumulated bonds)
    assert torch.equal(gen.nodes.cpu().to(torch.int8), torch.from_numpy(z["final_nodes"]))
    assert torch.equal(gen.edges.cpu().to(torch.int8), torch.from_numpy(z["final_edges"]))
    assert torch.equal(gen.n_nodes.cpu().to(torch.int8), torch.from_numpy(z["final_n_nodes"]))
    assert torch.equal(gen.likelihoods.cpu(), torch.from_numpy(z["final_likelihoods"]))


@pytest.mark.gpu
def test_sampling_with_the_pretrained_checkpoint_builds_molecules():
    path = pretrained_path()
    if path is None:
        pytest.skip("tests/golden/_local/pretrained_model.pth absent")
    from graphinvent_b200.config import make_constants
    from graphinvent_b200.generation import GraphGenerator
    from graphinvent_b200.gnn import mpnn
    C = make_constants("GGNN")
    net = mpnn.create(C)
    net.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
    net = net.cuda().eval()
    gen = GraphGenerator(net, batch_size=256, n_atom_types=5, n_formal_charge=3)
    g = torch.Generator(device="cuda").manual_seed(0)
    (nodes, edges, n_nodes), flat, final, proper = gen.sample(generator=g)
    assert nodes.shape == (256, 13, 8) and edges.shape == (256, 13, 13, 3)
    assert torch.isfinite(final).all() and (flat > 0).all()
    # same statistics as the reference run that made the trace (mean 10.4 atoms, 69 % properly terminated)
    assert 8.0 <= n_nodes.float().mean().item() <= 12.5
    assert proper.float().mean().item() >= 0.45
    atoms = (nodes.sum(-1) > 0).sum(-1)
    assert torch.equal(atoms.to(torch.int8), n_nodes)
    assert torch.equal(edges, edges.transpose(1, 2))
