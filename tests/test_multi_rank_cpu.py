"""world_size=2 over gloo on CPU: the host logic of the N>1 path (env sharding + the statistics all-reduce, the
engine's only collective).  The data path itself has no cross-rank step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fpyv_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard.env_shard(total, rank, world)
        # each rank "steps" its slice: 3 control steps, crashes on envs whose global index is a multiple of 1000
        idx = torch.arange(start, start + count)
        local = torch.zeros(8, dtype=torch.float64)
        local[0] = 3 * count
        local[1] = 3 * int((idx % 1000 == 0).sum())
        local[2] = local[1]
        local[3] = 5 * local[1]
        red = shard.reduce_stats(local)
        q.put((rank, start, count, red.tolist(), shard.stats_dict(red)["mean_episode_len"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [1 << 20, 1_000_003, 130])
def test_two_rank_sharding_and_stats_allreduce(total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, c0, red0, m0), (r1, s1, c1, red1, m1) = res
    assert s0 == 0 and s1 == c0 and c0 + c1 == total            # contiguous cover, no overlap
    assert abs(c0 - c1) <= 64
    crashes = 3 * len(range(0, total, 1000))
    assert red0 == red1 == [3.0 * total, crashes, crashes, 5.0 * crashes, 0, 0, 0, 0]
    assert m0 == m1 == 5.0


def test_env_shard_properties():
    for total in (0, 1, 63, 64, 65, 4096, 1 << 20, 16_777_216, 999_999):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.env_shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (s_a, c_a), (s_b, _) in zip(parts, parts[1:]):
                assert s_a + c_a == s_b
            if total % world == 0:
                assert len({c for _, c in parts}) == 1       # BASELINE configs divide evenly: equal slices
    with pytest.raises(ValueError):
        shard.env_shard(10, 2, 2)
    assert shard.env_shard(16_777_216, 7, 8) == (7 * 2_097_152, 2_097_152)


def test_reduce_stats_without_group_is_identity():
    t = torch.arange(8, dtype=torch.float64)
    assert torch.equal(shard.reduce_stats(t), t)
    assert shard.stats_dict(t)["episodes"] == 2.0
