"""
Data-parallel training of the hot path: molecules are independent units (no cross-graph term
anywhere in the model, `KLDivLoss(batchmean)` is a row sum / B), so the batch is split
contiguously across ranks, parameters are replicated, and the only exchange is ONE all-reduce
of the flat gradient bucket per step (NCCL over NVLink on the GPU box; gloo in the CPU tests).
The reference has no multi-GPU path; generation needs no communication at all.
"""
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """contiguous split, remainder spread over the first ranks: returns [lo, hi)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradAllReduce:
    """Callable installed as `model._grad_hook`: the fused backward hands it the flat fp32 gradient
    bucket (every parameter gradient is a view of it) right after the last backward kernel; it is
    summed across ranks in place and scaled by `local_batch / global_batch` weights so that unequal
    shards (last batch) still give the global batch-mean gradient."""

    def __init__(self, model, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.weight = 1.0 / self.world       # equal shards; set per step via set_shard() otherwise
        self.calls = 0
        self.bytes = 0
        model._grad_hook = self

    def set_shard(self, local_batch, global_batch):
        self.weight = float(local_batch) / float(global_batch)

    def __call__(self, flat):
        if self.world > 1:
            if self.weight != 1.0:
                flat.mul_(self.weight)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        return flat


def broadcast_parameters(model, src=0, group=None):
    """identical replicas at start (the optimizer then applies identical updates on every rank)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)
