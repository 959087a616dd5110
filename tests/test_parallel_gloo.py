"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path (contiguous batch
sharding + the single flat-bucket gradient all-reduce).  On the GPU box the same code runs over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphinvent_b200 import parallel


def test_shard_bounds_cover_the_batch_contiguously():
    for n in (0, 1, 7, 100, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(5))
        self._grad_hook = None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(1234)                       # same "dataset" on both ranks
        data = torch.randn(10, 5)
        model = _Tiny()
        with torch.no_grad():
            model.w.mul_(1.0 + rank)                  # deliberately different replicas ...
        parallel.broadcast_parameters(model)          # ... made identical from rank 0
        hook = parallel.GradAllReduce(model)
        lo, hi = parallel.shard_bounds(10, rank, world)   # 5 / 5
        # per-rank "batch mean" gradient bucket, as the fused backward hands it over
        flat = data[lo:hi].mean(0).clone()
        model._grad_hook(flat)
        want_equal = data.mean(0)
        # unequal shards (7 / 3): weights local/global reproduce the global batch mean
        lo2, hi2 = (0, 7) if rank == 0 else (7, 10)
        hook.set_shard(hi2 - lo2, 10)
        flat2 = data[lo2:hi2].mean(0).clone()
        model._grad_hook(flat2)
        # plain lists, not tensors: a tensor travels as a shared-memory handle that dies with this process
        q.put((rank, model.w.detach().tolist(), flat.tolist(), want_equal.tolist(), flat2.tolist(), hook.calls, hook.bytes))
    finally:
        dist.destroy_process_group()


def _run_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=120) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()          # exact PIDs we started
    if any(p.exitcode != 0 for p in procs):
        raise RuntimeError(f"worker exit codes {[p.exitcode for p in procs]}")
    return res


def test_flat_bucket_allreduce_world2_gloo():
    try:
        res = _run_world2()
    except Exception:             # the probed rendezvous port can be taken in between: one retry on a fresh port
        res = _run_world2()
    res = [(r, torch.tensor(w), torch.tensor(f), torch.tensor(wa), torch.tensor(f2), c, nb) for r, w, f, wa, f2, c, nb in res]
    for rank, w, flat, want, flat2, calls, nbytes in res:
        assert torch.equal(w, torch.ones(5))                      # broadcast from rank 0
        assert torch.allclose(flat, want, atol=1e-6)              # mean of equal shards == global mean
        assert torch.allclose(flat2, want, atol=1e-6)             # weighted unequal shards too
        assert calls == 2 and nbytes == 2 * 5 * 4
    assert torch.equal(res[0][2], res[1][2])                      # identical on both ranks
