"""CPU, world_size 2, gloo: the frame-sharding host logic (active_gs_b200/distributed.py) and the
sharded sampler agreement -- the N>1 path of the training loop without GPUs."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from active_gs_b200.distributed import FrameShard
    from active_gs_b200.gaussian_map import WeightedSampler
    from active_gs_b200.config import default_gaussian_map_config
    sh = FrameShard()
    assert sh.world == world and sh.rank == rank
    # same numpy seed on every rank -> same draw, disjoint contiguous slices that cover the batch
    cfg = default_gaussian_map_config().sampler
    cfg.batch_size = 8
    np.random.seed(77)
    sampler = WeightedSampler(cfg, 16)
    perf = torch.linspace(0.1, 1.6, 16)
    ids = sampler.next_ids(perf)
    mine = sh.my_frames(ids)
    assert sh.local_batch(len(ids)) == 4 and len(mine) == 4
    gathered = [None] * world
    dist.all_gather_object(gathered, (ids.tolist(), mine.tolist()))
    assert all(g[0] == ids.tolist() for g in gathered)
    assert sum((g[1] for g in gathered), []) == ids.tolist()
    # gradient all-reduce over views of one flat buffer = a single collective, summed over ranks
    flat = torch.arange(14 * 5, dtype=torch.float32) * (rank + 1)
    views, off = [], 0
    for n, shape in [(15, (5, 3)), (15, (5, 3)), (20, (5, 4)), (5, (5,)), (15, (5, 1, 3))]:
        views.append(flat[off:off + n].view(shape)); off += n
    sh.all_reduce_grads_(views)
    assert torch.equal(flat, torch.arange(70, dtype=torch.float32) * 3)
    assert torch.equal(views[2], (torch.arange(30, 50, dtype=torch.float32) * 3).view(5, 4))
    # visibility-count all-reduce (quirk Q1 coupling) and per-frame performance gather
    vis = torch.full((3, 4), rank + 1, dtype=torch.int32)
    assert int(sh.all_reduce_sum_(vis)[0, 0]) == 3
    pf = sh.gather_perf(torch.tensor([rank + 0.25, rank + 0.5]), None)
    assert torch.allclose(pf, torch.tensor([0.25, 0.5, 1.25, 1.5]))
    # a batch that does not divide by the world size is padded, not refused (the reference sampler
    # yields T < 8 frames for the first keyframes)
    assert sh.local_batch(7) == 4 and sh.pad(list(range(7))).tolist() == [0, 1, 2, 3, 4, 5, 6, -1]
    assert sh.my_frames(sh.pad([5, 6, 7])).tolist() == ([5, 6] if rank == 0 else [7, -1])
    # load-balanced partition: same answer on every rank, a permutation of the sampled ids, the active
    # keyframes at their pinned slots, and a smaller maximum load than the contiguous split
    cost = {int(i): float(1000 + 900 * (int(i) % 5)) for i in range(16) if i != 3}     # id 3 unknown -> mean
    bal = sh.balance(ids, len(sampler.active_ids), cost)
    dist.all_gather_object(gathered, bal.tolist())
    assert all(g == bal.tolist() for g in gathered)
    assert sorted(bal.tolist()) == sorted(ids.tolist())
    b = len(ids) // world
    for j, fid in enumerate(sampler.active_ids):
        r, k = sh.pinned_slot(j)
        assert bal[r * b + k] == fid
    mean = np.mean(list(cost.values()))
    load = lambda order: max(sum(cost.get(int(i), mean) for i in order[r * b:(r + 1) * b]) for r in range(world))
    assert load(bal) <= load(ids)
    assert sh.balance(ids, len(sampler.active_ids), {}).tolist() == sh.balance(ids, len(sampler.active_ids), {}).tolist()
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


def test_frame_shard_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))


def test_balance_partition_properties_all_world_sizes():
    """FrameShard.balance without a process group: permutation, B frames per rank, pinned active
    keyframes, never worse than the contiguous split by more than one frame's cost, and usually better."""
    from active_gs_b200.distributed import FrameShard
    rng = np.random.default_rng(0)
    better = worse = 0
    for world in (2, 3, 4, 8):
        for trial in range(40):
            b = int(rng.integers(1, 9))
            n_active = int(rng.integers(0, min(3, world * b) + 1))
            ids = rng.permutation(200)[:world * b]
            cost = {int(i): float(rng.uniform(2e4, 1.2e5)) for i in ids if rng.random() < 0.9}
            outs = []
            for rank in range(world):
                sh = FrameShard.__new__(FrameShard)
                sh.world, sh.rank = world, rank
                outs.append(sh.balance(ids, n_active, cost))
            assert all(np.array_equal(outs[0], o) for o in outs)          # rank-independent
            bal = outs[0]
            assert sorted(bal.tolist()) == sorted(ids.tolist()) and len(bal) == world * b
            for j in range(n_active):
                r, k = sh.pinned_slot(j)
                assert bal[r * b + k] == ids[j]
            mean = np.mean(list(cost.values())) if cost else 0.0
            c = lambda i: cost.get(int(i), mean)
            load = lambda order: max(sum(c(i) for i in order[r * b:(r + 1) * b]) for r in range(world))
            assert load(bal) <= load(ids) + max(c(i) for i in ids) + 1e-6
            better += load(bal) < load(ids) - 1e-6
            worse += load(bal) > load(ids) + 1e-6
    assert better > 100 and worse < 10


def test_padded_batches_for_the_first_keyframes_all_world_sizes():
    """with the reference sampler (batch_size 8, active_size 3) the sampled batch has
    v = T frames for the first 7 keyframes, so v % world != 0 is the normal case.  For T = 1..9 and world
    2/4/8: every rank gets the same number of slots, every real keyframe lands in exactly one slot, the
    active keyframes sit at their pinned slots, padding is -1."""
    from active_gs_b200.distributed import FrameShard
    from active_gs_b200.gaussian_map import WeightedSampler
    from active_gs_b200.config import default_gaussian_map_config
    cfg = default_gaussian_map_config().sampler
    rng = np.random.default_rng(1)
    for world in (2, 4, 8):
        for T in range(1, 10):
            np.random.seed(T)
            sampler = WeightedSampler(cfg, T)
            ids = sampler.next_ids(torch.linspace(0.5, 1.5, T))
            assert len(ids) == sampler.v == min(T, 8)
            cost = {int(i): float(rng.uniform(1e4, 9e4)) for i in ids}
            outs = []
            for rank in range(world):
                sh = FrameShard.__new__(FrameShard)
                sh.world, sh.rank = world, rank
                b = sh.local_batch(sampler.v)
                assert b == -(-sampler.v // world)
                bal = sh.balance(ids, len(sampler.active_ids), cost)
                assert len(bal) == b * world and len(sh.my_frames(bal)) == b
                outs.append(bal)
            assert all(np.array_equal(outs[0], o) for o in outs)
            real = outs[0][outs[0] >= 0]
            assert sorted(real.tolist()) == sorted(int(i) for i in ids)
            assert int((outs[0] < 0).sum()) == b * world - sampler.v
            for j, fid in enumerate(sampler.active_ids):
                r, k = sh.pinned_slot(j)
                assert outs[0][r * b + k] == fid
