"""CPU: host-side logic of the mirror that needs no kernel -- the capacity store behind the eight SoA map tensors
(active_gs_b200.gaussian_map._MapStore), the sampler's whole-population shortcut and the heap LPT partition
against a plain reference implementation."""
from types import SimpleNamespace

import numpy as np
import torch

from active_gs_b200 import ops
from active_gs_b200.distributed import FrameShard
from active_gs_b200.gaussian_map import WeightedSampler, _ATTR, _MapStore


def _fake_map(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    gm = SimpleNamespace()
    for name, w in ops.MAP_FIELDS:
        t = torch.rand((n, w) if w > 1 else (n,), generator=g)
        setattr(gm, _ATTR[name], t.view(n, 1, 3) if name == "harmonics" else t)
    return gm


def _snapshot(gm):
    return {name: getattr(gm, _ATTR[name]).clone() for name, _ in ops.MAP_FIELDS}


def _same(gm, snap):
    return all(torch.equal(getattr(gm, _ATTR[n]).reshape(snap[n].shape), snap[n]) for n in snap)


def test_map_store_adopts_grows_and_brings_external_tensors_home():
    st = _MapStore(torch.device("cpu"))
    gm = _fake_map(1000)
    snap = _snapshot(gm)
    st.adopt(gm, 500)                                   # first adoption: capacity buffers, both ping-pong halves
    assert st.owns(gm) and st.cap >= 2 * 1500 and st.alt is not None and _same(gm, snap)
    cap0, buf0 = st.cap, st.buf
    st.adopt(gm, 500)                                   # fits: nothing happens
    assert st.buf is buf0 and st.cap == cap0
    # the fused multi-GPU engine re-points the five parameter tensors at its symmetric flat buffer and trains there;
    # the three bookkeeping tensors stay views of the store.  adopt(gm, 0) must bring the parameters home WITHOUT a new
    # allocation and without clobbering the rows it is still reading (partly aliased: goes through the other half)
    ext = _fake_map(1000, seed=5)
    for name in ["means", "scales", "rotations", "opacities", "harmonics"]:
        setattr(gm, _ATTR[name], getattr(ext, _ATTR[name]))
    want = _snapshot(gm)
    assert not st.owns(gm)
    halves = {id(st.buf), id(st.alt)}
    st.adopt(gm, 0)
    assert st.owns(gm) and st.cap == cap0 and {id(st.buf), id(st.alt)} == halves and _same(gm, want)
    # growth keeps the rows and doubles
    st.adopt(gm, 10 * cap0)
    assert st.owns(gm) and st.cap >= 2 * (1000 + 10 * cap0) and _same(gm, want)
    # expose() after an in-place append / compaction: views of the first n rows, harmonics as (n,1,3)
    st.expose(gm, 1200)
    assert gm._means.shape == (1200, 3) and gm._harmonics.shape == (1200, 1, 3) and gm.view_means.shape == (1200, 3)
    assert gm._means.data_ptr() == st.buf["means"].data_ptr()


def test_sampler_takes_the_whole_population_without_a_draw():
    cfg = SimpleNamespace(active_size=3, batch_size=8)
    np.random.seed(3)
    state = np.random.get_state()[1].copy()
    s = WeightedSampler(cfg, 8)                          # 3 active + 5 others, batch 8: everything is selected
    ids = s.next_ids(np.full(8, 10.0, dtype=np.float32))
    assert sorted(ids.tolist()) == list(range(8)) and ids[:3].tolist() == [5, 6, 7]
    assert np.array_equal(np.random.get_state()[1], state), "the global numpy stream must not advance"
    s = WeightedSampler(cfg, 12)                         # 9 others, 5 drawn: the reference's np.random.choice
    np.random.seed(3)
    w = np.linspace(1.0, 2.0, 12).astype(np.float32)
    ids = s.next_ids(w)
    np.random.seed(3)
    p = w[:9] / np.sum(w[:9], dtype=np.float32)
    want = np.arange(9)[np.random.choice(np.arange(9), size=5, p=p, replace=False)]
    assert ids[:3].tolist() == [9, 10, 11] and ids[3:].tolist() == want.tolist()
    # torch tensors are accepted like numpy arrays (train_step passes a numpy view, tests a tensor)
    np.random.seed(3)
    assert s.next_ids(torch.from_numpy(w)).tolist() == ids.tolist()


def _reference_lpt(ids, n_active, cost, W, b):
    """the partition as first written: longest first onto the least loaded rank with a free slot"""
    ids = [int(i) for i in ids]
    known = [cost[i] for i in ids if i in cost]
    mean = float(np.mean(known)) if known else 0.0
    c = [float(cost.get(i, mean)) for i in ids]
    slots, load, free = [[None] * b for _ in range(W)], [0.0] * W, [b] * W
    for j in range(min(n_active, len(ids))):
        r, k = j % W, j // W
        slots[r][k] = ids[j]; load[r] += c[j]; free[r] -= 1
    for j in sorted(range(min(n_active, len(ids)), len(ids)), key=lambda j: (-c[j], j)):
        r = min((q for q in range(W) if free[q] > 0), key=lambda q: (load[q], q))
        slots[r][slots[r].index(None)] = ids[j]; load[r] += c[j]; free[r] -= 1
    return [(-1 if i is None else i) for r in range(W) for i in slots[r]]


def test_heap_partition_equals_the_plain_lpt():
    rng = np.random.default_rng(0)
    for W in (2, 3, 4, 8):
        sh = FrameShard.__new__(FrameShard)
        sh.world, sh.rank = W, 0
        for n in (1, 3, 7, 8, 15, 16, 31, 64):
            ids = rng.permutation(200)[:n]
            cost = {int(i): float(rng.integers(1, 50) * 100) for i in ids if rng.random() < 0.8}   # some unknown, many ties
            n_active = min(3, n)
            got = sh.balance(ids, n_active, cost).tolist()
            assert got == _reference_lpt(ids, n_active, cost, W, sh.local_batch(n)), (W, n)
