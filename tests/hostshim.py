"""
CPU stand-ins for a few libgib200.so entry points  --  TEST INFRASTRUCTURE for the host-logic tests (`-m "not gpu"`).

The product has no CPU path (CPU tensors raise; a missing library fails the import).  To exercise the Python host
logic around the kernels in this GPU-less container, a test can install these shims with pytest's monkeypatch:
`gib_generation_round` is played by the numpy oracle (`oracle/generation_oracle.py`) operating in place on the
caller's CPU tensors, `sample_actions` by torch's multinomial.  Nothing here is importable from the package.
"""
import ctypes
import types

import numpy as np
import torch

from oracle import generation_oracle as G


def _view(ptr, shape, ctype, dtype):
    addr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((ctype * n).from_address(addr)).view(dtype).reshape(shape)


def fake_generation_round(B, N, F, Ef, A, CH, rnd, action, lik, nodes, edges, n_nodes, likelihoods, g_nodes, g_edges,
                          g_n_nodes, g_lik, proper, cap, counters, scratch, stream):
    st = G.GenerationState.__new__(G.GenerationState)
    st.B, st.N, st.A, st.CH, st.Ef, st.F, st.rl = B, N, A, CH, Ef, F, False
    f32, i32, i8 = (ctypes.c_float, np.float32), (ctypes.c_int32, np.int32), (ctypes.c_int8, np.int8)
    st.nodes = _view(nodes, (B, N, F), *f32)
    st.edges = _view(edges, (B, N, N, Ef), *f32)
    st.n_nodes = _view(n_nodes, (B,), *i32)
    st.likelihoods = _view(likelihoods, (B, 2 * N), *f32)
    st.generated_nodes = _view(g_nodes, (cap, N, F), *f32)
    st.generated_edges = _view(g_edges, (cap, N, N, Ef), *f32)
    st.generated_n_nodes = _view(g_n_nodes, (cap,), *i8)
    st.generated_likelihoods = _view(g_lik, (cap, 2 * N), *f32)
    st.properly_terminated = _view(proper, (cap,), *i8)
    cnt = _view(counters, (2,), *i32)
    st.n_generated = int(cnt[0])
    written = G.generation_round(st, rnd, _view(action, (B,), *i32), _view(lik, (B,), *f32))
    cnt[0], cnt[1] = st.n_generated, written
    return 0


def fake_sample_actions(output, uniforms=None, generator=None):
    p = torch.softmax(output.detach(), dim=1)
    a = torch.multinomial(p, 1, generator=generator).squeeze(1)
    return a.to(torch.int32), p.gather(1, a.unsqueeze(1)).squeeze(1)


def install_generation_shims(monkeypatch):
    """route graphinvent_b200.generation's kernel calls to the CPU stand-ins (CPU tensors, device='cpu')"""
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200 import generation as gen
    fake = types.SimpleNamespace(gib_generation_round=fake_generation_round,
                                 gib_generation_scratch_bytes=lambda B: 64)
    monkeypatch.setattr(gen, "lib", fake)
    monkeypatch.setattr(Fn, "sample_actions", fake_sample_actions)
    monkeypatch.setattr(Fn, "build_graph", lambda model, edges: ("shared-graph", edges.data_ptr()))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))


class RecordingModelLib:
    """stand-in for the model entry points (`gib_graph_*`, `gib_model_*`): returns success / small sizes, computes
    nothing, and records the call sequence -- enough to test the argument plumbing, caching and autograd wiring of
    graphinvent_b200.functional on CPU tensors"""

    def __init__(self, params):
        self.calls = []
        self.numels = [p.numel() for p in params]

    def count(self, name):
        return self.calls.count(name)

    def __getattr__(self, name):
        if not name.startswith("gib_"):
            raise AttributeError(name)

        def entry(*args):
            self.calls.append(name)
            if name.endswith("_bytes"):
                return 256
            if name == "gib_model_num_params":
                return len(self.numels)
            if name == "gib_model_param_numel":
                return self.numels[args[1]]
            return 0
        return entry


def install_model_shims(monkeypatch, model):
    from graphinvent_b200 import functional as Fn
    fake = RecordingModelLib(list(model.parameters()))
    monkeypatch.setattr(Fn, "lib", fake)
    monkeypatch.setattr(Fn, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(Fn, "_stream", lambda device: None)
    return fake


def fake_adam_step(params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay, grad_scale, stream):
    """numpy restatement of torch/optim/adam.py::_single_tensor_adam on the flat buffers (fp32 like the kernel)"""
    f32 = np.float32
    P, Gr, M, V = (_view(x, (n,), ctypes.c_float, np.float32) for x in (params, grads, exp_avg, exp_avg_sq))
    g = (Gr * f32(grad_scale) + f32(weight_decay) * P).astype(f32)
    M[:] = M + (g - M) * f32(1.0 - beta1)
    V[:] = V * f32(beta2) + f32(1.0 - beta2) * g * g
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    P[:] = P - f32(lr / bc1) * (M / (np.sqrt(V) / f32(bc2 ** 0.5) + f32(eps)))
    return 0


def install_optim_shims(monkeypatch):
    from graphinvent_b200 import functional as Fn
    from graphinvent_b200 import optim
    calls = []

    def adam(*a):
        calls.append(a[4])           # n of every launch
        return fake_adam_step(*a)
    monkeypatch.setattr(optim, "lib", types.SimpleNamespace(gib_adam_step=adam))
    monkeypatch.setattr(Fn, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    return calls
