"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and gather in input order."""
import os
import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, str(ROOT))
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", ROOT / "ravvent_basecaller_b200" / "sharding.py")
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    ev = np.arange(n * 2, dtype=np.float32).reshape(n, 2)
    out = sh.run_sharded(lambda s: s[0].sum(axis=1, keepdims=True) + s[1].sum(axis=1, keepdims=True) + 0 * rank,
                         (raw, ev), rank, world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 1, 0, 10])
def test_two_rank_gloo_shard_and_gather(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    raw = np.arange(n * 3, dtype=np.float32).reshape(n, 3); ev = np.arange(n * 2, dtype=np.float32).reshape(n, 2)
    ref = raw.sum(axis=1, keepdims=True) + ev.sum(axis=1, keepdims=True)
    for r in range(2):
        assert np.array_equal(res[r].reshape(n, 1), ref)


def test_shard_range_partition():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", ROOT / "ravvent_basecaller_b200" / "sharding.py")
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    for n in (0, 1, 7, 8, 100000):
        for w in (1, 2, 4, 8):
            r = [sh.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
