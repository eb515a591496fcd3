"""Host-side multi-GPU logic on CPU: shard ranges and the episode-stat all-reduce over gloo (world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_simpletetris_b200.sharding import all_reduce_sum, shard_bounds


@pytest.mark.parametrize("n,w", [(4096, 1), (4096, 8), (262144, 4), (10, 3), (7, 8), (0, 2)])
def test_shard_bounds_partition(n, w):
    spans = [shard_bounds(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(n, w, w)


def test_all_reduce_identity_without_process_group():
    t = torch.tensor([1, 2, 3, 4])
    assert all_reduce_sum(t.clone()).tolist() == [1, 2, 3, 4]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(101, rank, world)
    stats = torch.tensor([hi - lo, 10 * (rank + 1), rank, 7], dtype=torch.int64)
    all_reduce_sum(stats)
    if rank == 0:
        torch.save(stats, out)
    dist.destroy_process_group()


def test_stats_all_reduce_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.load(out).tolist() == [101, 30, 1, 14]
