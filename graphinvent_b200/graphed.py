"""
One training step of the hot path as ONE CUDA-graph launch ("no host sync, static capacities,
CUDA-graph capturable").

    step = graphinvent_b200.graphed.TrainStep(model, optimizer, batch_size=B, entry_capacity=E_cap)
    for nodes, edges, target in loader:              # host (pinned) or device tensors, float32 or int8
        loss = step(nodes, edges, target)            # device scalar; float(loss) when the host wants it
        ...
    step.check()                                     # raises if a batch exceeded the capacity (one 64-byte read)

replaces the body of `Workflow.train_epoch` (Workflow.py:781-796: batch -> device, `model(nodes, edges)`, `loss`,
`zero_grad`, `backward`, `optimizer.step`) by: three copies into static input buffers, one graph launch
(K0 -> weight packing -> forward -> fused KL loss -> explicit backward into ONE flat gradient bucket), the
data-parallel all-reduce when a process group is given, and the optimizer step.  Capacity mode (functional.GraphBatch)
keeps every data-dependent extent in device memory, so the captured launch parameters never depend on the batch
content; a batch with more bond entries than `entry_capacity` is truncated and flagged, `check()` reports it.

The optimizer is stepped eagerly after the graph (its step count / learning-rate schedule are host state); with
`optim.FlatAdam` that is one more launch.  There is no CPU path.
"""
import ctypes

import torch

from . import functional as F
from ._lib import FLAG_MULTITYPE, FLAG_OVERFLOW, HDR_FLAGS, check, lib

_u8 = torch.uint8


class TrainStep:
    def __init__(self, model, optimizer, batch_size, entry_capacity, input_dtype=torch.float32, global_batch=None,
                 group=None, device=None, warmup=True):
        params = list(model.parameters())
        F._require_cuda(*params)
        self.model, self.optimizer = model, optimizer
        self.dev = device or params[0].device
        self.B = int(batch_size)
        self.global_batch = int(global_batch) if global_batch else self.B
        self.group = group
        self.code = 1 if input_dtype == torch.int8 else 0
        C = model.constants
        N, Fn, Ef = C.max_n_nodes, C.n_node_features, C.n_edge_features
        self.apd = N * (C.len_f_add_per_node + C.len_f_conn_per_node) + 1
        in_dt = torch.int8 if self.code else torch.float32
        dev = self.dev
        self.nodes = torch.zeros(self.B, N, Fn, dtype=in_dt, device=dev)
        self.edges = torch.zeros(self.B, N, N, Ef, dtype=in_dt, device=dev)
        self.target = torch.zeros(self.B, self.apd, dtype=torch.float32, device=dev)
        self.d = F.make_dims(model, self.B, self.code)
        d = self.d
        self.capacity = int(entry_capacity)
        # static buffers (addresses are baked into the graph)
        self.cws = torch.zeros(lib.gib_graph_count_ws_bytes(ctypes.byref(d)), dtype=_u8, device=dev)
        probe = F.GraphBatch(d, self.edges, capacity=self.capacity, buffers=None)
        self.cws, self.gbuf = probe.cws, probe.buf
        self.hdr_np, self.hdr = probe.hdr_np, probe.hdr
        F._check_params(model, d, params)
        self.params = params
        self.packed = torch.empty(lib.gib_model_packed_bytes(ctypes.byref(d)), dtype=_u8, device=dev)
        ws_bytes = lib.gib_model_workspace_bytes(ctypes.byref(d), self.hdr)
        if ws_bytes == 0:
            check(-1, "gib_model_workspace_bytes")
        self.ws = torch.empty(ws_bytes, dtype=_u8, device=dev)
        self.scratch = torch.empty(lib.gib_model_bwd_scratch_bytes(ctypes.byref(d), self.hdr), dtype=_u8, device=dev)
        self.out = torch.empty(self.B, self.apd, dtype=torch.float32, device=dev)
        self.dout = torch.empty_like(self.out)
        self.rows = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        total = sum(p.numel() for p in params)
        self.gflat = torch.zeros(total, dtype=torch.float32, device=dev)   # ONE bucket: grads are views of it
        self.views, o = [], 0
        for p in params:
            v = self.gflat[o:o + p.numel()].view(p.shape)
            self.views.append(v)
            p.grad = v
            o += p.numel()
        self.workspace_bytes = ws_bytes
        # data parallel: the gradients of the readout parameters (gather.*, APDReadout.*: the tail of the parameter
        # order, 79 % of the bucket) are final after the first part of the backward; their all-reduce runs on a side
        # stream while the message-passing backward (second captured graph) still executes
        self.world = 1
        if group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(group)
        self.tail_off = total
        if self.world > 1:
            names = [n for n, _ in model.named_parameters()]
            i = len(names)
            while i > 0 and names[i - 1].startswith(("gather.", "APDReadout.")):
                i -= 1
            self.tail_off = sum(p.numel() for p in params[:i])
            self.comm_stream = torch.cuda.Stream(dev)
        self.graph = None
        self.graph2 = None
        self.steps = 0
        self._param_ptrs = None
        if warmup:
            self.capture()

    # ---- the captured region ------------------------------------------------------------------------------
    def _enqueue(self):
        d, dev = self.d, self.dev
        st = F._stream(dev)
        bd = ctypes.byref(d)
        check(lib.gib_graph_count(bd, F._ptr(self.edges), F._ptr(self.cws), st), "gib_graph_count")
        check(lib.gib_graph_fill(bd, F._ptr(self.edges), F._ptr(self.cws), self.hdr, F._ptr(self.gbuf), st),
              "gib_graph_fill")
        check(lib.gib_model_pack(bd, F._ptr_table(self.params), F._ptr(self.packed), st), "gib_model_pack")
        check(lib.gib_model_forward(bd, self.hdr, F._ptr(self.nodes), F._ptr(self.edges), F._ptr(self.gbuf),
                                    F._ptr(self.packed), F._ptr(self.ws), F._ptr(self.out), st), "gib_model_forward")
        # Workflow.loss (Workflow.py:833-860) with the batch-mean taken over the GLOBAL batch (data-parallel shards)
        check(lib.gib_kl_loss_fwd_bwd(F._ptr(self.out), F._ptr(self.target), self.B, self.apd,
                                      1.0 / self.global_batch, F._ptr(self.rows), F._ptr(self.dout), st),
              "gib_kl_loss_fwd_bwd")
        check(lib.gib_sum_scaled(F._ptr(self.rows), self.B, 1.0 / self.global_batch, F._ptr(self.loss), st),
              "gib_sum_scaled")
        check(lib.gib_fill_zero(F._ptr(self.gflat), self.gflat.numel() * 4, st), "gib_fill_zero")
        self._backward(1 if self.world > 1 else 0)

    def _backward(self, part):
        st = F._stream(self.dev)
        check(lib.gib_model_backward_part(ctypes.byref(self.d), self.hdr, F._ptr(self.nodes), F._ptr(self.edges),
                                          F._ptr(self.gbuf), F._ptr(self.packed), F._ptr(self.ws), F._ptr(self.out),
                                          F._ptr(self.dout), F._ptr_table(self.views), F._ptr(self.scratch), part, st),
              "gib_model_backward_part")

    def _enqueue_all(self):
        """the whole step's launch sequence, eagerly (warm-up, kernel-class timing)"""
        self._enqueue()
        if self.world > 1:
            self._backward(2)

    def capture(self):
        """(re)capture; called by the constructor and again if the parameters were moved (e.g. by FlatAdam)"""
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):            # warm-up outside capture: lazy per-device init, function attributes
            self._enqueue_all()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue()
        self.graph = g
        if self.world > 1:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self._backward(2)
            self.graph2 = g2
        self._param_ptrs = [p.data_ptr() for p in self.params]

    # ---- one step -----------------------------------------------------------------------------------------
    def load(self, nodes, edges, target):
        """copy one batch into the static input buffers (pinned host tensors: asynchronous H2D)"""
        if nodes.shape[0] != self.B:
            raise ValueError(f"TrainStep was built for batches of {self.B} molecules, got {nodes.shape[0]}")
        self.nodes.copy_(nodes, non_blocking=True)
        self.edges.copy_(edges, non_blocking=True)
        self.target.copy_(target, non_blocking=True)

    def __call__(self, nodes=None, edges=None, target=None):
        if nodes is not None:
            self.load(nodes, edges, target)
        if [p.data_ptr() for p in self.params] != self._param_ptrs:
            self.capture()                        # the parameters moved (optimizer re-flattened them)
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v                        # zero_grad(set_to_none=True) of the reference loop (Workflow.py:787)
        self.graph.replay()
        if self.world > 1:
            # per-rank gradients are already scaled by 1/global_batch: a plain sum is the global batch mean
            dist, cur = torch.distributed, torch.cuda.current_stream(self.dev)
            grp = self.group if self.group not in (None, False) else None
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):          # readout gradients: overlapped with graph 2
                if self.tail_off < self.gflat.numel():
                    dist.all_reduce(self.gflat[self.tail_off:], op=dist.ReduceOp.SUM, group=grp)
            self.graph2.replay()                               # message-passing backward
            if self.tail_off > 0:
                dist.all_reduce(self.gflat[:self.tail_off], op=dist.ReduceOp.SUM, group=grp)
            cur.wait_stream(self.comm_stream)
        self.optimizer.step()
        F.invalidate_packed_weights()
        self.steps += 1
        return self.loss

    def check(self):
        """synchronising read of the K0 flags of the LAST step; raises if it did not fit the capacity"""
        flags = int(self.cws[: 64].view(torch.int32).cpu()[HDR_FLAGS])
        if flags & FLAG_OVERFLOW:
            raise RuntimeError(f"a batch held more bond entries than entry_capacity={self.capacity}; "
                               "the results of that step are invalid -- rebuild TrainStep with a larger capacity")
        if flags & FLAG_MULTITYPE and self.d.model == F.MODEL_ID["AttGGNN"]:
            raise RuntimeError("AttentionGGNN requires one bond type per bond (as the reference's AggregationMPNN does)")
        return flags
