"""GPU: every kernel family through the C-ABI against the ATen expression it replaces."""
import ctypes

import pytest
import torch

from tests.conftest import load_small

pytestmark = pytest.mark.gpu


def _env():
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200._lib import lib, check
    return Fn, lib, check


def _p(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad16(x):
    return (x + 15) // 16 * 16


def _padded(t, ld):
    out = torch.zeros(t.shape[0], ld, device="cuda")
    out[:, : t.shape[1]] = t
    return out


# ------------------------------------------------------------------------------------------
# K0: bond-entry lists + CSR vs the reference's nonzero() ordering
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["GGNN", "EMN"])
@pytest.mark.parametrize("case", ["small", "c2ish", "wide"])
def test_graph_build_matches_nonzero(model, case):
    Fn, lib, check = _env()
    from graphinvent_b200 import synthetic as S
    from graphinvent_b200.gnn import mpnn
    from oracle import mpnn_oracle as O
    if case == "small":
        fx = load_small(model)
        C, edges = fx["C"], fx["edges"]
    elif case == "c2ish":
        C = O.make_constants(model)
        _, e = S.random_graphs(300, 13, 5, 3, seed=4, min_atoms=0)
        edges = torch.from_numpy(e).float()
    else:
        C = O.make_constants(model, max_n_nodes=40, n_node_features=12, len_f_add_per_node=81)
        _, e = S.random_graphs(64, 40, 9, 3, seed=5)
        edges = torch.from_numpy(e).float()
        edges[3, 0, 0, 1] = 1.0            # a self loop
        edges[5, 2, 7, :] = torch.tensor([0.5, 0.0, 2.0])  # multi-type, non-binary bond
        edges[5, 7, 2, :] = torch.tensor([0.5, 0.0, 2.0])
    net = mpnn.create(C)
    B, N, Ef = edges.shape[0], edges.shape[1], edges.shape[3]
    d = Fn.make_dims(net, B)
    g = Fn.GraphBatch(d, edges.cuda())
    torch.cuda.synchronize()
    by_type = model != "EMN"
    G = Ef if by_type else 1
    hdr = g.hdr_np
    S_ = B * N
    if by_type:
        nz = (edges != 0)
        E = int(nz.sum())
    else:
        nz = (edges != 0).any(-1, keepdim=True)
        E = int(nz.sum())
    assert hdr[0] == E
    P = int(hdr[1])
    ent_src = g.array(d, 0, P).cpu(); ent_dst = g.array(d, 1, P).cpu()
    ent_w = g.array(d, 2, P, torch.float32).cpu()
    dst_ptr = g.array(d, 3, S_ + 1).cpu(); dst_ent = g.array(d, 4, E).cpu()
    src_ptr = g.array(d, 5, S_ + 1).cpu(); src_ent = g.array(d, 6, E).cpu()
    base = 0
    for t in range(G):
        b, i, j = nz[..., t].nonzero(as_tuple=True)      # row-major (b, i, j): the reference order
        cnt = b.numel()
        assert hdr[2 + t] == cnt and hdr[6 + t] == base and base % 128 == 0
        sl = slice(base, base + cnt)
        assert torch.equal(ent_dst[sl].long(), b * N + i)
        assert torch.equal(ent_src[sl].long(), b * N + j)
        want_w = edges[b, i, j, t] if by_type else torch.ones(cnt)
        assert torch.equal(ent_w[sl], want_w)
        pad_to = base + (cnt + 127) // 128 * 128
        assert (ent_src[base + cnt: pad_to] == -1).all() and (ent_w[base + cnt: pad_to] == 0).all()
        base = pad_to
    assert hdr[6 + G] == base == P
    # CSR by destination: entries ordered (b, i, j, t); by source: (b, j, i, t)
    b, i, j, t = nz.nonzero(as_tuple=True)
    deg = torch.zeros(S_, dtype=torch.long).index_add_(0, b * N + i, torch.ones_like(b))
    assert torch.equal(dst_ptr.long(), torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]))
    assert torch.equal(ent_dst[dst_ent.long()].long(), b * N + i)
    assert torch.equal(ent_src[dst_ent.long()].long(), b * N + j)
    assert len(set(dst_ent.tolist())) == E
    key = ((b * N + j) * N + i) * G + t
    order = torch.argsort(key)
    sdeg = torch.zeros(S_, dtype=torch.long).index_add_(0, b * N + j, torch.ones_like(b))
    assert torch.equal(src_ptr.long(), torch.cat([torch.zeros(1, dtype=torch.long), sdeg.cumsum(0)]))
    assert torch.equal(ent_src[src_ent.long()].long(), (b * N + j)[order])
    assert torch.equal(ent_dst[src_ent.long()].long(), (b * N + i)[order])
    flags = int(hdr[11])
    assert bool(flags & 1) == bool(((edges != 0).sum(-1) > 1).any())
    assert bool(flags & 2) == bool(((edges != 0) & (edges != 1)).any())


# ------------------------------------------------------------------------------------------
# dense kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 16, 16), (37, 45, 100), (300, 250, 250), (5000, 500, 136),
                                   (20000, 250, 250), (20000, 3, 500), (1024, 585, 500), (129, 1, 685)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_fwd(M, N, K, act):
    Fn, lib, check = _env()
    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    Kp, Np = pad16(K), pad16(N)
    Xp, Wp = _padded(X, Kp), torch.zeros(Np, Kp, device="cuda")
    Wp[:N, :K] = W
    bp = torch.zeros(Np, device="cuda"); bp[:N] = b
    Y = torch.full((M, Np), float("nan"), device="cuda")
    check(lib.gib_linear_fwd(_p(Xp), Kp, _p(Wp), Kp, _p(bp), _p(Y), Np, M, Np, Kp, act, _st()), "linear_fwd")
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    ref = {0: ref, 1: torch.selu(ref), 2: torch.tanh(ref)}[act]
    assert torch.isfinite(Y).all()
    assert (Y[:, :N].double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert (Y[:, N:] == 0).all()          # pad columns are exact zeros (layout contract)


@pytest.mark.parametrize("M,N,K", [(1, 16, 16), (100, 45, 100), (5000, 250, 250), (40000, 500, 500),
                                   (1024, 585, 685), (70, 300, 128)])
def test_linear_bwd_dw(M, N, K):
    Fn, lib, check = _env()
    torch.manual_seed(M + N)
    G = torch.randn(M, N, device="cuda")
    X = torch.randn(M, K, device="cuda")
    Np, Kp = pad16(N), pad16(K)
    Gp, Xp = _padded(G, Np), _padded(X, Kp)
    dW = torch.ones(N, K, device="cuda")        # accumulate semantics: starts at 1
    db = torch.ones(N, device="cuda")
    sc = torch.empty(lib.gib_dw_scratch_bytes(M, Np, Kp), dtype=torch.uint8, device="cuda")
    check(lib.gib_linear_bwd_dw(_p(Gp), Np, Np, _p(Xp), Kp, Kp, M, _p(dW), _p(db), N, K, _p(sc), None, None, _st()),
          "dw")
    ref = G.double().t() @ X.double() + 1
    refb = G.double().sum(0) + 1
    # fp32 accumulation of M unit-variance products (entries ~ sqrt(M), max ~ 5 sqrt(M)); the tcgen05 path adds the
    # tensor core's truncating accumulate: measured <= 1e-5 of the largest entry
    tol = 2e-5 * ref.abs().max().item() + 1e-5
    assert (dW.double() - ref).abs().max().item() <= tol
    assert (db.double() - refb).abs().max().item() <= 2e-5 * refb.abs().max().item() + 1e-4


def _planes(lib, check, W):
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    check(lib.gib_split_planes(_p(W), _p(hi), _p(lo), W.numel(), _st()), "split_planes")
    return hi, lo


@pytest.mark.parametrize("gen", [2, 1])
@pytest.mark.parametrize("M,N,K,act", [(128, 128, 32, 0), (100, 48, 16, 1), (1000, 256, 256, 1), (23808, 256, 128, 1),
                                       (5000, 608, 512, 0), (13312, 256, 144, 1), (1024, 500, 688, 1)])
def test_tcgen05_linear_planes_both_generations(gen, M, N, K, act):
    """the model's call pattern: pre-split weight planes, activation split inside the kernel -- second generation
    (operand through tensor memory) and first generation (through shared memory) against fp64"""
    Fn, lib, check = _env()
    torch.manual_seed(M + N + K)
    X = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    hi, lo = _planes(lib, check, W)
    assert (hi.double() + lo.double() - W.double()).abs().max().item() <= 2.0 ** -21 * W.abs().max().item()   # two 11-bit pieces
    Y = torch.full((M, N), float("nan"), device="cuda")
    lib.gib_tc_debug(1 if gen == 1 else 0)
    try:
        check(lib.gib_linear_fwd_tc_planes(_p(X), K, _p(hi), _p(lo), K, _p(b), _p(Y), N, M, N, K, act, None, None,
                                           _st()), "linear_fwd_tc_planes")
        torch.cuda.synchronize()
    finally:
        lib.gib_tc_debug(0)
    ref = torch.nn.functional.linear(X.double(), W.double(), b.double())
    ref = torch.selu(ref) if act else ref
    assert torch.isfinite(Y).all()
    assert (Y.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


def test_tcgen05_device_side_row_ranges():
    """capacity mode: the live row range of a problem is read from device memory (K0's bond-type group sizes); rows
    outside it may hold anything and must stay untouched"""
    Fn, lib, check = _env()
    torch.manual_seed(5)
    cap, N, K = 4096, 256, 256
    X = torch.randn(cap, K, device="cuda")
    X[3000:] = float("nan")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    hi, lo = _planes(lib, check, W)
    G = torch.randn(cap, N, device="cuda")
    G[3000:] = float("nan")
    sc = torch.empty(lib.gib_dw_scratch_bytes(cap, N, K), dtype=torch.uint8, device="cuda")
    for base, m in [(0, 1000), (128, 2500), (2944, 56), (0, 0)]:
        md = torch.tensor([m], dtype=torch.int32, device="cuda")
        bd = torch.tensor([base], dtype=torch.int32, device="cuda")
        Y = torch.full((cap, N), 7.0, device="cuda")
        check(lib.gib_linear_fwd_tc_planes(_p(X), K, _p(hi), _p(lo), K, _p(b), _p(Y), N, cap, N, K, 1, _p(md), _p(bd),
                                           _st()), "linear dyn")
        ref = torch.selu(torch.nn.functional.linear(X[base:base + m].double(), W.double(), b.double()))
        if m:
            assert (Y[base:base + m].double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
        assert (Y[:base] == 7.0).all() and (Y[base + m:] == 7.0).all()
        dW = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
        check(lib.gib_linear_bwd_dw(_p(G), N, N, _p(X), K, K, cap, _p(dW), _p(db), N, K, _p(sc), _p(md), _p(bd),
                                    _st()), "dw dyn")
        rw = G[base:base + m].double().t() @ X[base:base + m].double()
        rb = G[base:base + m].double().sum(0)
        assert (dW.double() - rw).abs().max().item() <= 2e-5 * rw.abs().max().item() + 1e-5
        assert (db.double() - rb).abs().max().item() <= 2e-5 * rb.abs().max().item() + 1e-4


def _random_csr(S, E, seed):
    g = torch.Generator().manual_seed(seed)
    dst = torch.randint(0, S, (E,), generator=g).sort().values
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    ptr[1:] = torch.bincount(dst, minlength=S).cumsum(0).int()
    ent = torch.randperm(E, generator=g).int()        # arbitrary entry rows
    return dst, ptr, ent


@pytest.mark.parametrize("S,E,width", [(50, 0, 16), (1000, 2400, 100), (13312, 28672, 128), (155648, 352256, 100)])
def test_scatter_sum_is_the_dense_summation_matmul(S, E, width):
    Fn, lib, check = _env()
    ld = pad16(width)
    dst, ptr, ent = _random_csr(S, E, S + E)
    msg = torch.randn(max(E, 1), ld, device="cuda")
    w = torch.rand(max(E, 1), device="cuda") + 0.5
    out = torch.full((S, ld), float("nan"), device="cuda")
    ptr_d, ent_d = ptr.cuda(), ent.cuda()          # keep the device copies alive across the async launch
    check(lib.gib_scatter_sum(_p(out), _p(msg), ld, _p(ptr_d), _p(ent_d), _p(w), S, _st()), "scatter")
    ref = torch.zeros(S, ld, dtype=torch.float64, device="cuda")
    if E:
        rows = ent.long().cuda()
        ref.index_add_(0, dst.cuda(), (msg[rows] * w[rows, None]).double())
    assert (out.double() - ref).abs().max().item() <= 1e-5
    if E and S * E <= 5_000_000:   # the literal reference formulation: dense [S,E] 0/1 matrix (summation_mpnn.py:111-141)
        dense = (torch.arange(S)[:, None] == dst[None, :]).float().cuda()
        ref2 = dense @ (msg[ent.long().cuda()] * w[ent.long().cuda(), None])
        assert (out - ref2).abs().max().item() <= 1e-4


def test_seg_softmax_matches_padded_softmax():
    Fn, lib, check = _env()
    S, E, ld = 700, 1500, 48
    dst, ptr, ent = _random_csr(S, E, 3)
    EM = torch.randn(E, ld, device="cuda"); EN = torch.randn(E, ld, device="cuda") * 3
    w = torch.ones(E, device="cuda")
    out = torch.empty(S, ld, device="cuda")
    ptr_d, ent_d = ptr.cuda(), ent.cuda()
    check(lib.gib_seg_softmax(_p(out), _p(EM), _p(EN), ld, _p(ptr_d), _p(ent_d), _p(w), S, _st()), "seg")
    ref = torch.zeros(S, ld, dtype=torch.float64)
    EMc, ENc = EM.cpu().double(), EN.cpu().double()
    for s in range(S):
        rows = ent[ptr[s]:ptr[s + 1]].long()
        if rows.numel():
            ref[s] = (torch.softmax(ENc[rows], 0) * EMc[rows]).sum(0)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-5


def test_gru_gates_match_grucell():
    Fn, lib, check = _env()
    torch.manual_seed(0)
    S, H, Min = 500, 100, 84
    Hp = pad16(H)
    cell = torch.nn.GRUCell(Min, H).cuda()
    x = torch.randn(S, Min, device="cuda"); h = torch.randn(S, H, device="cuda")
    gi = torch.nn.functional.linear(x, cell.weight_ih, cell.bias_ih)
    gh = torch.nn.functional.linear(h, cell.weight_hh, cell.bias_hh)
    def blocked(g):
        o = torch.zeros(S, 3 * Hp, device="cuda")
        for k in range(3):
            o[:, k * Hp: k * Hp + H] = g[:, k * H:(k + 1) * H]
        return o
    ptr = torch.zeros(S + 1, dtype=torch.int32)
    active = torch.rand(S) < 0.7
    ptr[1:] = active.int().cumsum(0).int()
    hn = torch.empty(S, Hp, device="cuda")
    gi_b, gh_b, h_p, ptr_d = blocked(gi), blocked(gh), _padded(h, Hp), ptr.cuda()
    check(lib.gib_gru_gates(_p(hn), _p(gi_b), _p(gh_b), _p(h_p), Hp, _p(ptr_d), S, _st()), "gru")
    ref = cell(x, h)
    ref = torch.where(active.cuda()[:, None], ref, h)
    assert (hn[:, :H] - ref).abs().max().item() <= 2e-6
    assert (hn[:, H:] == 0).all()


def test_graph_gather_reproduces_masked_softmax_quantisation():
    """modules.py:44-52 incl. the fp32 `energies - 1e6` arithmetic for molecules without a bonded atom."""
    Fn, lib, check = _env()
    torch.manual_seed(1)
    B, N, W = 40, 13, 100
    ld = pad16(W)
    en = torch.randn(B * N, W, device="cuda") * 2; em = torch.randn(B * N, W, device="cuda")
    mask = torch.rand(B, N) < 0.6
    mask[0] = False; mask[1] = False; mask[2] = True          # fully masked molecules and a full one
    ptr = torch.zeros(B * N + 1, dtype=torch.int32)
    ptr[1:] = mask.view(-1).int().cumsum(0).int()
    g = torch.empty(B, ld, device="cuda"); att = torch.empty(B * N, ld, device="cuda")
    en_p, em_p, ptr_d = _padded(en, ld), _padded(em, ld), ptr.cuda()
    check(lib.gib_graph_gather(_p(g), _p(att), _p(en_p), _p(em_p), ld, _p(ptr_d), N, B, 1e6, _st()), "gather")
    energies = en.view(B, N, W) - ((mask == 0).float() * 1e6).cuda().unsqueeze(-1)
    ref = (torch.softmax(energies, dim=1) * em.view(B, N, W)).sum(1)
    assert (g[:, :W] - ref).abs().max().item() <= 2e-6
    assert (g[:, W:] == 0).all()


def test_kl_loss_and_sampling():
    from graphinvent_b200 import functional as Fn
    from oracle import mpnn_oracle as O
    torch.manual_seed(2)
    out = (torch.randn(64, 625, device="cuda") * 2).requires_grad_(True)
    tgt = torch.rand(64, 625, device="cuda")
    tgt[:, ::3] = 0
    loss = Fn.kl_loss(out, tgt)
    loss.backward()
    o2 = out.detach().cpu().requires_grad_(True)
    ref = O.kl_loss(o2, tgt.cpu())
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5
    assert (out.grad.cpu() - o2.grad).abs().max().item() <= 1e-7
    # inverse-CDF sampling: index decode must agree with a float64 cumulative sum of the same softmax
    u = torch.rand(64, device="cuda")
    a, lik = Fn.sample_actions(out.detach(), uniforms=u)
    p = torch.softmax(out.detach().double(), 1)
    cdf = p.cumsum(1)
    lo = torch.where(a > 0, cdf.gather(1, (a.long() - 1).clamp(min=0)[:, None])[:, 0], torch.zeros_like(u).double())
    hi = cdf.gather(1, a.long()[:, None])[:, 0]
    assert ((lo - 1e-5 <= u.double()) & (u.double() <= hi + 1e-5)).all()
    assert (lik.double() - p.gather(1, a.long()[:, None])[:, 0]).abs().max().item() <= 1e-6
