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


@pytest.mark.gpu
def test_sharded_basecaller_matches_single_device():
    """One process, all visible GPUs (2 under `gpurun --gpus 2`; a single device exercises the same threads / staging
    path): the sharded call returns, in input order, exactly what one device returns for the whole batch."""
    import torch
    import ravvent_basecaller_b200 as rb
    from oracle import model_ref as mr
    w = mr.init_weights(22, random_bias=True)
    n = 1500
    x = mr.synth_chunks(np.random.default_rng(3), n)
    devs = list(range(torch.cuda.device_count()))
    sb = rb.ShardedBasecaller(128, 128, 128, rb.nuc_tk, "joint", 0., devices=devs, wave_snippets=256).load_weights(w)
    one = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., device=devs[-1], wave_snippets=256).load_weights(w)
    for W in (1, 5):
        ids, sc = sb.beam_search_prediction(x, W, 14)
        rid, rsc = one.beam_search_prediction(x, W, 14)
        assert ids.shape == rid.shape and np.array_equal(ids, rid) and np.array_equal(sc, rsc)
    # ragged: fewer items than devices
    ids, sc = sb.beam_search_prediction((x[0][:1], x[1][:1]), 5, 14)
    rid, rsc = one.beam_search_prediction((x[0][:1], x[1][:1]), 5, 14)
    assert np.array_equal(ids, rid) and np.array_equal(sc, rsc)


@pytest.mark.gpu
def test_host_pipeline_pageable_equals_pinned_equals_device():
    """rvb_beam_host: pageable numpy buffers (pinned staging ring), page-locked buffers (direct copies) and device
    tensors give identical results over several waves, including a ragged last wave."""
    import torch
    import ravvent_basecaller_b200 as rb
    from oracle import model_ref as mr
    w = mr.init_weights(22, random_bias=True)
    n = 5 * 192 + 17
    raw, ev = mr.synth_chunks(np.random.default_rng(4), n)
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., wave_snippets=192).load_weights(w)
    a_ids, a_sc = bc.beam_search_prediction((raw, ev), 5, 12)                            # pageable
    pr, pe = torch.from_numpy(raw).pin_memory(), torch.from_numpy(ev).pin_memory()
    b_ids, b_sc = bc.beam_search_prediction((pr.numpy(), pe.numpy()), 5, 12)             # pinned
    d_ids, d_sc = bc.beam_search_prediction((pr.cuda(), pe.cuda()), 5, 12)               # device
    assert np.array_equal(a_ids, b_ids) and np.array_equal(a_sc, b_sc)
    assert np.array_equal(a_ids, d_ids.cpu().numpy()) and np.array_equal(a_sc, d_sc.cpu().numpy())
    with pytest.raises(ValueError):
        bc.beam_search_prediction((raw, ev[:, :, :4]), 5, 12)
    with pytest.raises(ValueError):
        bc.beam_search_prediction((raw, ev[:-1]), 5, 12)
    bc.load_weights(w)                                                                   # second finalize on one handle: no leak, same results
    c_ids, _ = bc.beam_search_prediction((raw, ev), 5, 12)
    assert np.array_equal(a_ids, c_ids)
