"""The N > 1 path on CPU: two gloo ranks each take their contiguous shard (the device is stood in for by
the oracle, which is allowed in tests), results are gathered and must equal the single-process result;
the barrier / max-over-ranks timing reduction used by bench.py is exercised too."""
import os
import random
import sys

import numpy as np
import pytest

from common import ROOT, cases_to_arrays, signature_cases

sys.path.insert(0, os.path.join(ROOT, "babyjubjub-rs_b200"))


def test_shard_ranges_partition_the_batch():
    from babyjubjub_rs_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from babyjubjub_rs_b200.sharding import max_over_ranks, shard_range
    import common
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = common.signature_cases(random.Random(3), 3)
    arrs = common.cases_to_arrays(cases)
    n = len(cases)
    lo, hi = shard_range(n, rank, world)
    ora = common.OracleC(threads=1)
    ok = ora.verify(*[a[lo:hi] for a in arrs])          # the per-rank "device" pass
    dist.barrier()
    full = torch.zeros(n, dtype=torch.uint8)
    full[lo:hi] = torch.from_numpy(ok)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)         # test-side gather only; the product has no collective
    t = max_over_ranks(1.0 + rank, dist)
    if rank == 0:
        np.save(os.path.join(tmp, "ok.npy"), full.numpy())
        np.save(os.path.join(tmp, "t.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_verify(tmp_path, oracle_c):
    import torch.multiprocessing as mp
    port = 29500 + random.Random(os.getpid()).randrange(2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cases = signature_cases(random.Random(3), 3)
    exp = oracle_c.verify(*cases_to_arrays(cases))
    assert np.array_equal(np.load(tmp_path / "ok.npy"), exp)
    assert float(np.load(tmp_path / "t.npy")[0]) == 2.0      # max over ranks of (1 + rank)


def test_multigpu_helper_runs_every_shard_once():
    from babyjubjub_rs_b200.sharding import MultiGpu
    seen = []
    mg = MultiGpu(lambda d: "engine%d" % d, [0, 1, 2])
    mg.run_sharded(10, lambda e, lo, hi: seen.append((e, lo, hi)))
    assert sorted(seen) == [("engine0", 0, 3), ("engine1", 3, 6), ("engine2", 6, 10)]
    with pytest.raises(RuntimeError):
        mg.run_sharded(3, lambda e, lo, hi: (_ for _ in ()).throw(RuntimeError("boom")))
