"""
Parameter containers with the reference's constructor signatures and state_dict keys
(reference `gnn/modules.py`: `MLP` :111-170, `GraphGather` :12-52, `GlobalReadout` :173-281).

In this implementation the arithmetic of these blocks is fused into the whole-model CUDA path
(`graphinvent_b200/csrc/model.cu`), so the containers only own parameters: calling one on its
own raises -- there is deliberately no ATen fallback.
"""
import torch


class _FusedOnly(torch.nn.Module):
    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} is a parameter container: its computation is fused into the "
            "model-level sm_100a kernels (call the owning GGNN/MNN/AttentionGGNN/EMN module).")


class MLP(_FusedOnly):
    """Linear -> SELU -> AlphaDropout per layer, activation after the last layer too.
    Linear layers sit at `seq.0, seq.3, ...` exactly as in the reference (modules.py:133-142)."""

    def __init__(self, in_features: int, hidden_layer_sizes: list, out_features: int, dropout_p: float) -> None:
        super().__init__()
        self.dropout_p = float(dropout_p)
        fs = [in_features, *hidden_layer_sizes, out_features]
        layers = []
        for in_f, out_f in zip(fs, fs[1:]):
            linear = torch.nn.Linear(in_f, out_f, bias=True)
            torch.nn.init.xavier_uniform_(linear.weight)       # modules.py:162-163
            layers += [linear, torch.nn.SELU(), torch.nn.AlphaDropout(dropout_p)]
        self.seq = torch.nn.Sequential(*layers)


class GraphGather(_FusedOnly):
    def __init__(self, node_features: int, hidden_node_features: int, out_features: int, att_depth: int,
                 att_hidden_dim: int, att_dropout_p: float, emb_depth: int, emb_hidden_dim: int,
                 emb_dropout_p: float, big_positive: float) -> None:
        super().__init__()
        self.big_positive = big_positive
        self.att_nn = MLP(node_features + hidden_node_features, [att_hidden_dim] * att_depth, out_features,
                          att_dropout_p)
        self.emb_nn = MLP(hidden_node_features, [emb_hidden_dim] * emb_depth, out_features, emb_dropout_p)


class GlobalReadout(_FusedOnly):
    def __init__(self, f_add_elems: int, f_conn_elems: int, f_term_elems: int, mlp1_depth: int,
                 mlp1_dropout_p: float, mlp1_hidden_dim: int, mlp2_depth: int, mlp2_dropout_p: float,
                 mlp2_hidden_dim: int, graph_emb_size: int, max_n_nodes: int, node_emb_size: int,
                 device: str) -> None:
        super().__init__()
        self.device = device
        self.fAddNet1 = MLP(node_emb_size, [mlp1_hidden_dim] * mlp1_depth, f_add_elems, mlp1_dropout_p)
        self.fConnNet1 = MLP(node_emb_size, [mlp1_hidden_dim] * mlp1_depth, f_conn_elems, mlp1_dropout_p)
        self.fAddNet2 = MLP(max_n_nodes * f_add_elems + graph_emb_size, [mlp2_hidden_dim] * mlp2_depth,
                            f_add_elems * max_n_nodes, mlp2_dropout_p)
        self.fConnNet2 = MLP(max_n_nodes * f_conn_elems + graph_emb_size, [mlp2_hidden_dim] * mlp2_depth,
                             f_conn_elems * max_n_nodes, mlp2_dropout_p)
        self.fTermNet2 = MLP(graph_emb_size, [mlp2_hidden_dim] * mlp2_depth, f_term_elems, mlp2_dropout_p)
