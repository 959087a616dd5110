"""
CPU oracle for one round of the reference's batched graph generator  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Plain numpy restatement of `GraphGenerator.build_graphs`'s per-round tensor work (reference
`GraphGenerator.py:118-157`): `get_actions` decode (:504-570), `get_invalid_actions` (:573-657),
`copy_terminated_graphs` (:340-385), `apply_actions` (:211-338), `reset_graphs` (:425-465), for the 6-tuple action
layout (atom type + formal charge node features; no implicit-H / chirality segment).  The sampled flat APD index and
its likelihood are INPUTS (the reference draws them with torch's Multinomial), so the state machine can be replayed.

Parity pin: `tests/golden/generation_trace.npz` is a recording of the unmodified reference generator (draws of every
round + its final buffers, made by `tests/golden/make_generation_trace.py`); `tests/test_generation.py` replays the
draws through this oracle and requires bit-exact buffers.  The CUDA kernels are then checked against this oracle on
random action streams as well.

RL twin (`GraphGeneratorRL.py:109-172, 227-438, 520-629`): the same state machine carrying TWO likelihood streams
(agent = the sampling model, prior = the second model's probability of the same action); pass `rl=True` /
`prior_likelihoods=`.  Pinned by `tests/golden/generation_rl_trace.npz` (`make_generation_rl_trace.py`).
"""
import numpy as np


class GenerationState:
    """the tensors `GraphGenerator` keeps (initialize_graph_batch :387-423, allocate_graph_tensors :163-209)"""

    def __init__(self, batch, N, A, CH, Ef, rl=False):
        self.B, self.N, self.A, self.CH, self.Ef, self.F = batch, N, A, CH, Ef, A + CH
        self.rl = rl
        self.nodes = np.zeros((batch, N, self.F), np.float32)
        self.edges = np.zeros((batch, N, N, Ef), np.float32)
        self.n_nodes = np.zeros(batch, np.int32)
        self.likelihoods = np.zeros((batch, 2 * N), np.float32)
        self.nodes[0] = 1                       # the dummy graph in slot 0 (:418-423)
        self.edges[0, 0, 0, 0] = 1
        self.n_nodes[0] = 1
        cap = 2 * batch
        self.generated_nodes = np.zeros((cap, N, self.F), np.float32)
        self.generated_edges = np.zeros((cap, N, N, Ef), np.float32)
        self.generated_n_nodes = np.zeros(cap, np.int8)
        self.generated_likelihoods = np.zeros((cap, 2 * N), np.float32)
        self.properly_terminated = np.zeros(cap, np.int8)
        self.n_generated = 0
        if rl:                                  # GraphGeneratorRL.allocate_graph_tensors :200-217
            self.prior_likelihoods = np.zeros((batch, 2 * N), np.float32)
            self.generated_prior_likelihoods = np.zeros((cap, 2 * N), np.float32)


def decode(state, b, a):
    """flat APD index -> (kind, bond_to, bond_from, atom, charge, bond_type, invalid)   kind: 0 add, 1 connect, 2 terminate"""
    N, A, CH, Ef = state.N, state.A, state.CH, state.Ef
    n = int(state.n_nodes[b])
    len_add, len_conn = N * A * CH * Ef, N * Ef
    if a < len_add:                                           # f_add[bond_to, atom, charge, bond_type]   (:504-516)
        bt, at, ch, ty = np.unravel_index(a, (N, A, CH, Ef))
        bf = n                                                # :557
        empty = n == 0
        invalid = ((not empty) and bt >= n) or (empty and bt != 0) or bf >= N     # :600-613
        if bf >= N or empty:                                  # :568 `f_add_idc[5][max_node_idc] = 0`
            bf = 0
        return 0, int(bt), int(bf), int(at), int(ch), int(ty), bool(invalid)
    if a < len_add + len_conn:                                # f_conn[bond_to, bond_type]
        bt, ty = np.unravel_index(a - len_add, (N, Ef))
        bf = n - 1                                            # :561
        bfw = bf + N if bf < 0 else bf                        # Python negative indexing of the reference tensors
        invalid = bt >= n or n == 0 or bt == bf or state.edges[b, bt, bfw].sum() == 1   # :616-629
        return 1, int(bt), int(bfw), 0, 0, int(ty), bool(invalid)
    return 2, 0, 0, 0, 0, 0, False


def generation_round(state, rnd, actions, likelihoods, prior_likelihoods=None):
    """one pass of the `while` body of build_graphs (:118-157; RL: GraphGeneratorRL.py:127-168); returns the number
    of graphs written this round"""
    B, A = state.B, state.A
    rl = prior_likelihoods is not None
    assert rl == state.rl
    rec = [decode(state, b, int(actions[b])) for b in range(B)]
    term = [b for b in range(B) if rec[b][0] == 2]
    invalid = [b for b in range(B) if rec[b][6]]
    k = state.n_generated
    cap = state.properly_terminated.shape[0]
    state.properly_terminated[k:min(cap, k + len(term))] = 1                        # :127 (counts slot 0 too)
    order = [b for b in term if b != 0] + [b for b in invalid if b != 0]          # :130-133
    for i, b in enumerate(order):                                                 # copy_terminated_graphs
        state.likelihoods[b, rnd] = likelihoods[b]
        if rl:
            state.prior_likelihoods[b, rnd] = prior_likelihoods[b]
        p = k + i
        if p < cap:
            state.generated_nodes[p] = state.nodes[b]
            state.generated_edges[p] = state.edges[b]
            state.generated_n_nodes[p] = state.n_nodes[b]
            state.generated_likelihoods[p] = state.likelihoods[b]
            if rl:
                state.generated_prior_likelihoods[p] = state.prior_likelihoods[b]
    state.n_generated = k + len(order)
    gone = set(order)
    for b in range(B):                                                            # apply_actions on every slot
        kind, bt, bf, at, ch, ty, _ = rec[b]
        if b in gone:
            continue                                                              # reset below anyway
        if kind == 0:
            state.nodes[b, bf, at] = 1
            state.nodes[b, bf, A + ch] = 1
            if state.n_nodes[b] != 0:
                state.edges[b, bt, bf, ty] = 1
                state.edges[b, bf, bt, ty] = 1
            state.n_nodes[b] += 1
        elif kind == 1:
            state.edges[b, bf, bt, ty] = 1
            state.edges[b, bt, bf, ty] = 1
        if kind in (0, 1):
            state.likelihoods[b, rnd] = likelihoods[b]
            if rl:
                state.prior_likelihoods[b, rnd] = prior_likelihoods[b]
    for b in order:                                                               # reset_graphs
        state.nodes[b] = 0
        state.edges[b] = 0
        state.n_nodes[b] = 0
        state.likelihoods[b] = 0
        if rl:
            state.prior_likelihoods[b] = 0
    state.nodes[0] = 1                                                            # dummy graph re-stamped (:462-465)
    state.edges[0, 0, 0, 0] = 1
    state.n_nodes[0] = 1
    return len(order)
