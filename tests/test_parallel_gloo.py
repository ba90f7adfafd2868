"""CPU: the N>1 plumbing (utterance sharding + the single gather of token ids) with
world_size 2 over gloo.  The data path has no other collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qwen3_asr_rs_b200 import parallel


def test_shard_range_covers_everything_once():
    for n in [1, 2, 3, 7, 8, 9, 64, 129]:
        for world in [1, 2, 3, 4, 8]:
            seen = []
            for r in range(world):
                lo, hi = parallel.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _fake_ids(i, max_new):
    n = (i * 7) % (max_new + 1)
    return [1000 * i + k for k in range(n)]


def _worker(rank, world, port, n_items, max_new, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n_items, world, rank)
        local = [_fake_ids(i, max_new) for i in range(lo, hi)]
        out = parallel.gather_token_ids(local, n_items, max_new)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [2, 5])
def test_gather_token_ids_world2(n_items):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    max_new = 9
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, max_new, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [_fake_ids(i, max_new) for i in range(n_items)]
    assert res[0] == want and res[1] == want
