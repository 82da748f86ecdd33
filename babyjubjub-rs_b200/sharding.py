"""Multi-GPU plumbing.  The path shards with no exchange between lanes, so this is only:
contiguous shard ranges, a barrier, and the max-over-ranks time (torch.distributed, NCCL on GPUs /
gloo in the CPU tests).  No data-path collective exists by design."""
import threading


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one, empty shards allowed"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (identity without a process group)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class MultiGpu:
    """One Engine per visible device, one host thread per device, contiguous shards, no NCCL."""

    def __init__(self, engine_factory, devices):
        self.engines = [engine_factory(d) for d in devices]

    def run_sharded(self, n, fn):
        """fn(engine, lo, hi) is called once per device on its own thread; exceptions propagate"""
        g = len(self.engines)
        errs = [None] * g

        def work(d):
            lo, hi = shard_range(n, d, g)
            try:
                if hi > lo:
                    fn(self.engines[d], lo, hi)
            except Exception as e:      # noqa: BLE001
                errs[d] = e
        th = [threading.Thread(target=work, args=(d,)) for d in range(g)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in errs:
            if e is not None:
                raise e
