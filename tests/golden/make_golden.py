"""
Generates the golden fixtures in this directory by running the UNMODIFIED reference
(`/root/reference/graphinvent/gnn`, imported per SURVEY.md Appendix C) on CPU fp32.
Run in the build container (the reference is not mounted on the GPU box):

    python tests/golden/make_golden.py

Outputs
  small_<MODEL>.npz   tiny-dims config per model: constants (json), reference-initialised
                      state_dict, int8 inputs (random molecules + the generator's corner
                      graphs), targets, reference logits / loss / parameter gradients.
  gdb13_rows.npz      first 256 real rows of data/pre-training/gdb13_1K/train.h5 (int8).
  gdb13_pretrained_logits.npz
                      reference logits + loss + per-tensor gradient max/rms for those rows
                      through the shipped GGNN checkpoint
                      (data/fine-tuning/gdb13_1K-debug/pretrained_model.pth).
  _local/pretrained_model.pth
                      byte copy of that checkpoint (23.7 MB of DATA, git-ignored; it
                      travels to the GPU box with the gpurun snapshot).
"""
import hashlib
import json
import os
import shutil
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from graphinvent_b200 import synthetic as S          # noqa: E402
from oracle import mpnn_oracle as O                  # noqa: E402  (make_constants / kl_loss only)
from tests import refimpl                            # noqa: E402

SMALL = dict(n_node_features=6, n_edge_features=3, max_n_nodes=7, len_f_add_per_node=9,
             len_f_conn_per_node=3, hidden_node_features=12, message_size=10, message_passes=3,
             enn_hidden_dim=20, enn_depth=2, msg_hidden_dim=20, msg_depth=2, att_hidden_dim=18,
             att_depth=2, gather_width=11, gather_att_hidden_dim=16, gather_att_depth=2,
             gather_emb_hidden_dim=14, gather_emb_depth=2, mlp1_hidden_dim=24, mlp1_depth=2,
             mlp2_hidden_dim=28, mlp2_depth=2, edge_emb_size=12, edge_emb_hidden_dim=20,
             edge_emb_depth=2)


def small_constants(model):
    kw = dict(SMALL)
    if model == "MNN":
        kw["message_size"] = 12
    return O.make_constants(model, **kw)


def constants_json(C):
    return json.dumps({k: getattr(C, k) for k in C._fields})


def run_reference(net, nodes, edges, target):
    out = net(nodes, edges)
    loss = O.kl_loss(out, target)          # restated Workflow.py:833-860 (Workflow is not importable)
    net.zero_grad()
    loss.backward()
    grads = OrderedDict((k, p.grad.detach().clone()) for k, p in net.named_parameters())
    return out.detach(), loss.detach(), grads


def make_small(model, seed):
    torch.manual_seed(seed)
    C = small_constants(model)
    net = refimpl.build(C)
    n1, e1 = S.random_graphs(27, C.max_n_nodes, 4, 2, seed=seed + 10, min_atoms=0)
    n2, e2 = S.corner_case_graphs(C.max_n_nodes, C.n_node_features)
    nodes, edges = np.concatenate([n2, n1]), np.concatenate([e2, e1])
    apd = C.max_n_nodes * (C.len_f_add_per_node + C.len_f_conn_per_node) + 1
    target = S.random_targets(nodes.shape[0], apd, seed=seed)
    out, loss, grads = run_reference(net, torch.from_numpy(nodes).float(),
                                     torch.from_numpy(edges).float(), torch.from_numpy(target))
    blob = {"constants": np.array(constants_json(C)), "nodes": nodes, "edges": edges,
            "target": target, "logits": out.numpy(), "loss": loss.numpy()}
    for k, v in net.state_dict().items():
        blob["param/" + k] = v.numpy()
    for k, v in grads.items():
        blob["grad/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"small_{model}.npz"), **blob)
    print(model, "rows", nodes.shape[0], "loss", float(loss))


def make_gdb13():
    ref_data = "/root/reference/data"
    nodes, edges, apds = refimpl.read_gdb13_h5(f"{ref_data}/pre-training/gdb13_1K/train.h5")
    nodes, edges, apds = nodes[:256].copy(), edges[:256].copy(), apds[:256].copy()
    np.savez_compressed(os.path.join(HERE, "gdb13_rows.npz"), nodes=nodes, edges=edges, apds=apds)
    os.makedirs(os.path.join(HERE, "_local"), exist_ok=True)
    src = f"{ref_data}/fine-tuning/gdb13_1K-debug/pretrained_model.pth"
    dst = os.path.join(HERE, "_local", "pretrained_model.pth")
    shutil.copyfile(src, dst)
    sha = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    C = O.make_constants("GGNN")
    net = refimpl.build(C)
    net.load_state_dict(torch.load(dst, map_location="cpu", weights_only=False))
    out, loss, grads = run_reference(net, torch.from_numpy(nodes).float(),
                                     torch.from_numpy(edges).float(),
                                     torch.from_numpy(apds).float())
    blob = {"logits": out.numpy(), "loss": loss.numpy(), "sha256": np.array(sha),
            "grad_names": np.array(list(grads.keys())),
            "grad_absmax": np.array([float(g.abs().max()) for g in grads.values()], np.float32),
            "grad_rms": np.array([float(g.pow(2).mean().sqrt()) for g in grads.values()], np.float32),
            "grad_sum": np.array([float(g.double().sum()) for g in grads.values()], np.float64)}
    # full gradients of a few small / sensitive tensors
    for k in ("gru.bias_ih", "gru.bias_hh", "msg_nns.2.seq.12.bias", "APDReadout.fTermNet2.seq.12.weight",
              "gather.att_nn.seq.0.bias", "msg_nns.0.seq.0.bias"):
        blob["grad/" + k] = grads[k].numpy()
    np.savez_compressed(os.path.join(HERE, "gdb13_pretrained_logits.npz"), **blob)
    print("gdb13 pretrained: loss", float(loss), "sha256", sha[:16])


if __name__ == "__main__":
    assert refimpl.available(), "needs /root/reference (build container only)"
    torch.set_num_threads(8)
    for i, m in enumerate(("GGNN", "MNN", "AttGGNN", "EMN")):
        make_small(m, 100 + i)
    make_gdb13()
