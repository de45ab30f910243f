"""World-size-2 gloo tests (CPU) of the view-parallel host logic: camera sharding, the single gradient all-reduce and the
partition-independent significance reduction.  The rasterizer itself has no CPU path, so a deterministic stand-in
count_render is used here; the GPU version of the same check is tests/test_gpu_multi.py."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightgaussian_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_count_render(cam, gaussians, pipe, bg):
    g = torch.Generator().manual_seed(1000 + cam)
    P = gaussians.get_xyz.shape[0]
    return {"gaussians_count": torch.randint(0, 500, (P,), generator=g, dtype=torch.int32)}


def _model(P=257):
    g = torch.Generator().manual_seed(7)
    return SimpleNamespace(get_xyz=torch.rand(P, 3, generator=g), get_opacity=torch.rand(P, 1, generator=g))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # 1. significance: sharded == serial, bit for bit
    model = _model()
    cams = list(range(11))
    cnt, imp = parallel.sharded_prune_list(model, cams, None, None, _fake_count_render, rank, world)
    # 2. gradients: one all-reduce over the flat buffer == sum of the per-rank gradients
    params = [torch.zeros(5, 3), torch.zeros(7), torch.zeros(2, 2, 2)]
    flat = parallel.FlatGrads(params)
    for i, p in enumerate(params):
        p.grad.add_(torch.full_like(p, float((rank + 1) * (i + 1))))
    flat.allreduce(world)
    torch.save(dict(cnt=cnt, imp=imp, grads=[p.grad.clone() for p in params]), os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_matches_serial(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    model = _model()
    serial_cnt, serial_imp = parallel.sharded_prune_list(model, list(range(11)), None, None, _fake_count_render, 0, 1)
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(d["cnt"], serial_cnt)              # integer sums commute: identical for any partition
        assert torch.equal(d["imp"], serial_imp)              # score = opacity * total count, same bits on every rank
        for i, g in enumerate(d["grads"]):
            assert torch.equal(g, torch.full_like(g, float((1 + 2) * (i + 1))))


def test_shard_views_partitions_all_cameras():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in parallel.shard_views(37, r, world))
        assert seen == list(range(37))
    assert parallel.shard_views(5, 7, 8) == []


def test_flat_grads_views_alias_the_buffer():
    params = [torch.zeros(4, 3, requires_grad=True), torch.zeros(6, requires_grad=True)]
    flat = parallel.FlatGrads(params)
    (params[0].sum() * 2 + params[1].sum() * 3).backward()
    assert torch.equal(flat.flat, torch.cat([torch.full((12,), 2.0), torch.full((6,), 3.0)]))
    flat.zero()
    assert params[0].grad.abs().sum() == 0 and flat.nbytes == 18 * 4


def test_balanced_view_schedule_groups_views_of_similar_cost():
    g = torch.Generator().manual_seed(3)
    costs = (torch.rand(16, generator=g) * 10 + 5).tolist()
    for world in (2, 4, 8):
        sched = parallel.balanced_view_schedule(costs, world)
        assert sorted(i for step in sched for i in step) == list(range(16))          # every view exactly once
        assert all(len(step) == world for step in sched)
        spread = max(max(costs[i] for i in s) - min(costs[i] for i in s) for s in sched)
        naive = max(max(costs[(t * world + r) % 16] for r in range(world)) - min(costs[(t * world + r) % 16] for r in range(world))
                    for t in range(16 // world))
        assert spread <= naive
        # the slowest rank of every step, summed: never worse than the round-robin assignment
        assert sum(max(costs[i] for i in s) for s in sched) <= sum(max(costs[(t * world + r) % 16] for r in range(world))
                                                                   for t in range(16 // world)) + 1e-9
    # an incomplete last group repeats its own views; no rank is left without work
    sched = parallel.balanced_view_schedule([3.0, 1.0, 2.0, 5.0, 4.0], 4)
    assert len(sched) == 2 and all(len(s) == 4 for s in sched) and set(sched[1]) == {1}
    assert parallel.balanced_view_schedule(costs, 2) == parallel.balanced_view_schedule(list(costs), 2)   # deterministic
