"""
Hyper-parameter tuple for the drop-in modules when no reference `parameters.constants` is
around (benchmarks, tests, standalone use).  Field names and default values are the
reference's (`parameters/defaults.py:93-128, 145-433`; derived dims `constants.py:158-211`);
any namedtuple/object with these attributes works -- the reference's own `constants` does.
"""
from collections import namedtuple

DEFAULTS = dict(
    model="GGNN", device="cuda", big_positive=1e6, big_negative=-1e6,
    n_node_features=8, n_edge_features=3, max_n_nodes=13,          # gdb13: 5 atom types + 3 charges
    len_f_add_per_node=45, len_f_conn_per_node=3,
    hidden_node_features=100, message_size=100, message_passes=3,
    enn_hidden_dim=250, enn_depth=4, enn_dropout_p=0.0,
    msg_hidden_dim=250, msg_depth=4, msg_dropout_p=0.0,
    att_hidden_dim=250, att_depth=4, att_dropout_p=0.0,
    gather_width=100,
    gather_att_hidden_dim=250, gather_att_depth=4, gather_att_dropout_p=0.0,
    gather_emb_hidden_dim=250, gather_emb_depth=4, gather_emb_dropout_p=0.0,
    mlp1_hidden_dim=500, mlp1_depth=4, mlp1_dropout_p=0.0,
    mlp2_hidden_dim=500, mlp2_depth=4, mlp2_dropout_p=0.0,
    edge_emb_size=100, edge_emb_hidden_dim=250, edge_emb_depth=4, edge_emb_dropout_p=0.0,
)


def make_constants(model="GGNN", **overrides):
    d = dict(DEFAULTS, model=model)
    d.update(overrides)
    d.setdefault("edge_features", d["n_edge_features"])          # names EdgeMPNN.__init__ reads
    d.setdefault("edge_embedding_size", d["edge_emb_size"])      # (edge_mpnn.py:16-17)
    return namedtuple("constants", sorted(d))(**d)


def apd_length(C):
    return C.max_n_nodes * (C.len_f_add_per_node + C.len_f_conn_per_node) + 1
