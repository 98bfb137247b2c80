"""GPU parity of the per-keyframe map maintenance kernels:
ags_spawn, ags_view_stats_update, ags_prune_compact, ags_view_utility -- each against the CPU oracle
(oracle/host_ref.py, pinned to the reference's own Python by tests/test_oracle_spawn.py) on the same
inputs, through the C ABI."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from active_gs_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden", "spawn_golden.pt")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _store(cap, dev):
    from active_gs_b200 import ops
    o = dict(device=dev, dtype=torch.float32)
    return {n: (torch.full((cap, w), 7.0, **o) if w > 1 else torch.full((cap,), 7.0, **o)) for n, w in ops.MAP_FIELDS}


def _spawn(frame, pred, dev, voxel_size, seed=0, n_old=0, cap=None, smooth=None):
    from active_gs_b200 import ops
    rgb, depth = frame["rgb"].to(dev).contiguous(), frame["depth"].to(dev).contiguous()
    _, H, W = rgb.shape
    smooth = ops.smooth_depth(depth) if smooth is None else smooth.to(dev).contiguous()
    cap = cap or (n_old + H * W)
    st = _store(cap, dev)
    sel = torch.zeros(H * W, dtype=torch.uint8, device=dev)
    p = None if pred is None else tuple(t.to(dev).contiguous() for t in pred)
    n_new, cand, wanted = ops.spawn(rgb, depth, smooth, frame["extrinsic"].float(), torch.linalg.inv(frame["intrinsic"].float()),
                                    p, st, n_old, cap, error_thres=0.25, voxel_size=voxel_size, seed=seed, select_out=sel)
    return st, n_new, cand, wanted, sel.cpu(), smooth.cpu()


def _assert_matches(st, n_old, n_new, new):
    assert n_new == new["means"].shape[0]
    sl = slice(n_old, n_old + n_new)
    assert torch.allclose(st["means"][sl].cpu(), new["means"], atol=2e-6)
    assert torch.allclose(st["rotations"][sl].cpu(), new["rotations"], atol=2e-5)
    assert torch.equal(st["harmonics"][sl].cpu(), new["harmonics"][:, 0, :])
    assert torch.equal(st["scales"][sl].cpu(), new["scales"])
    for k in ["opacities", "view_scores", "view_supports"]:
        assert float(st[k][sl].abs().max()) == 0.0
    assert float(st["view_means"][sl].abs().max()) == 0.0
    for k in st:                                                     # rows outside the appended range untouched
        assert bool((st[k][:n_old] == 7.0).all()) and bool((st[k][n_old + n_new:] == 7.0).all())


def test_spawn_first_keyframe_matches_reference_fixture(gold):
    """no map yet: every valid, front-facing pixel spawns (voxel filter off = the fixture's 'keep all')"""
    from oracle import host_ref as hr
    dev = _dev()
    g = gold["first"]
    sm_ref = hr.smooth_depth(g["frame"]["depth"])
    # (a) the spawn kernel alone: same smoothed depth on both sides
    st, n_new, cand, wanted, sel, _ = _spawn(g["frame"], None, dev, voxel_size=0.0, n_old=5, smooth=sm_ref)
    assert cand == n_new == wanted
    _assert_matches(st, 5, n_new, g["new"])
    ref = hr.spawn_candidates(g["frame"])
    assert torch.equal(sel > 0, ref["select"])
    # (b) the whole device chain (CUDA bilateral filter -> spawn): the normals amplify the filter's
    # 1e-5 rounding differences by 1 / pixel spacing
    st, n_new, _, _, sel, smooth = _spawn(g["frame"], None, dev, voxel_size=0.0)
    assert torch.allclose(smooth, sm_ref, atol=1e-5)
    assert torch.equal(sel > 0, ref["select"])
    assert torch.allclose(st["rotations"][:n_new].cpu(), g["new"]["rotations"], atol=2e-3)


def test_spawn_on_initialised_map_matches_reference_fixture(gold):
    """cal_mask against a render of the map; the oracle's render is fed to both sides"""
    from oracle import host_ref as hr, rasterizer_ref as rr
    dev = _dev()
    g = gold["second"]
    s, f = g["state"], g["frame"]
    attrs = hr.activate(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"],
                        s["view_scores"], s["view_supports"], s["view_means"])
    with torch.no_grad():
        out = hr.render_view_all(rr.rasterize, f["extrinsic"][None], f["intrinsic"][None], attrs, torch.zeros(4),
                                 (0.001, 10.0), g["hw"])
    st, n_new, cand, wanted, sel, _ = _spawn(f, (out[0][0], out[1][0], out[3][0]), dev, voxel_size=0.0,
                                             smooth=hr.smooth_depth(f["depth"]))
    _assert_matches(st, 0, n_new, g["new"])


def _plane_frame(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    K = syn.normalised_intrinsic(H, W, 60.0)
    ext = torch.eye(4)
    ext[:3, 3] = torch.tensor([0.3, -0.2, 0.1])
    xs = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
    ys = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    depth = (1.0 + 0.002 * xs + 0.001 * ys)[None].clone()
    depth[torch.rand(1, H, W, generator=g) < 0.02] = -1.0
    return dict(rgb=torch.rand(3, H, W, generator=g), depth=depth, extrinsic=ext, intrinsic=K,
                depth_range=torch.tensor([0.0, 5.0]))


def test_spawn_voxel_filter_one_per_voxel_random_and_reproducible():
    """the hash-voxel filter keeps exactly one candidate per occupied 2 cm voxel (what
    voxel_downsample guarantees, utils/operations.py:603-625), in pixel order; the choice depends on
    the seed only"""
    from oracle import host_ref as hr
    dev = _dev()
    f = _plane_frame(120, 160, 3)
    st0, n_all, cand, _, sel0, _ = _spawn(f, None, dev, voxel_size=0.0)
    points = st0["means"][:n_all].cpu()
    runs = []
    for seed in (11, 11, 12):
        st, n_new, cand2, wanted, sel, _ = _spawn(f, None, dev, voxel_size=0.02, seed=seed)
        assert cand2 == cand == n_all and wanted == n_new
        assert torch.equal(sel > 0, sel0 > 0)
        picked = torch.nonzero((sel[sel0 > 0] == 2)).flatten()       # indices into the candidate list
        assert picked.numel() == n_new < n_all
        assert hr.voxel_filter_is_valid(points, picked, 0.02)
        assert torch.equal(st["means"][:n_new].cpu(), points[picked])
        runs.append(picked)
    assert torch.equal(runs[0], runs[1]) and not torch.equal(runs[0], runs[2])
    # the winners are spread over the members of a voxel, not always the first / last one
    vox = hr.voxel_ids(points, 0.02)
    _, inv = torch.unique(vox, dim=0, return_inverse=True)
    first_of_voxel = torch.full((int(inv.max()) + 1,), n_all, dtype=torch.long).scatter_reduce(0, inv, torch.arange(n_all), "amin")
    frac_first = (first_of_voxel[inv[runs[0]]] == runs[0]).float().mean()
    assert 0.1 < float(frac_first) < 0.9


def test_spawn_capacity_clamp_and_errors():
    from active_gs_b200 import ops
    dev = _dev()
    f = _plane_frame(32, 48, 4)
    st, n_new, cand, wanted, sel, _ = _spawn(f, None, dev, voxel_size=0.0, n_old=10, cap=10 + 100)
    assert n_new == 100 and wanted == cand > 100
    assert bool((st["means"][:10] == 7.0).all())
    with pytest.raises(RuntimeError):
        ops.spawn(f["rgb"].to(dev), f["depth"].to(dev), f["depth"].to(dev), torch.eye(4), torch.eye(3), None,
                  _store(8, dev), 9, 8, error_thres=0.25)


def test_view_stats_update_matches_oracle():
    from active_gs_b200 import ops
    from oracle import host_ref as hr
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    N = 5000
    state = syn.make_room_scene(N, seed=21)
    state["view_supports"] = torch.randint(0, 4, (N,), generator=g).float()
    state["view_means"] = torch.randn(N, 3, generator=g) * 0.3
    state["view_scores"] = torch.rand(N, generator=g)
    counts = torch.randint(0, 3, (N,), generator=g, dtype=torch.int32)
    cam = torch.tensor([0.4, -0.3, 1.4])
    for use_vd in (True, False):
        ref = {k: v.clone() for k, v in state.items()}
        hr.view_stats_update(ref, counts >= 1, cam, 5.0, use_vd)
        d = {k: state[k].clone().to(dev) for k in ["means", "rotations", "view_supports", "view_means", "view_scores"]}
        ops.view_stats_update(counts.to(dev), d["means"], d["rotations"], cam, 5.0, use_vd, d["view_supports"],
                              d["view_means"], d["view_scores"])
        for k in ["view_supports", "view_means", "view_scores"]:
            assert torch.allclose(d[k].cpu(), ref[k], atol=1e-6), k


@pytest.mark.parametrize("N", [1, 255, 256, 70001])
def test_prune_compact_matches_boolean_indexing(N):
    from active_gs_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(N)
    src = {n: (torch.randn(N, w, generator=g) if w > 1 else torch.randn(N, generator=g) * 3) for n, w in ops.MAP_FIELDS}
    counts = (torch.rand(3, N, generator=g) < 0.3).to(torch.int32)
    mask = torch.rand(N, generator=g) < 0.2
    drop = mask | (counts.sum(0) < 1) | (torch.sigmoid(src["opacities"]) < 0.1)
    sd = {k: v.to(dev) for k, v in src.items()}
    dd = {k: torch.full_like(v, -5.0) for k, v in sd.items()}
    md = mask.clone().to(dev)
    n = ops.prune_compact(sd, dd, N, counts=counts.to(dev), prune_mask=md)
    assert n == int((~drop).sum())
    assert torch.equal(md.cpu(), drop)                                # quirk Q5: caller's mask updated in place
    for k in src:
        assert torch.equal(dd[k][:n].cpu(), src[k][~drop]), k
        assert bool((dd[k][n:] == -5.0).all())
    # mask-only form (GaussianMap.prune) and the no-mask form
    md2 = mask.clone().to(dev)
    n2 = ops.prune_compact(sd, dd, N, prune_mask=md2)
    drop2 = mask | (torch.sigmoid(src["opacities"]) < 0.1)
    assert n2 == int((~drop2).sum()) and torch.equal(dd["rotations"][:n2].cpu(), src["rotations"][~drop2])
    n3 = ops.prune_compact(sd, dd, N)
    assert n3 == int((torch.sigmoid(src["opacities"]) >= 0.1).sum())


def test_view_utility_matches_oracle(gold):
    from active_gs_b200 import ops, operations as O
    from oracle import host_ref as hr
    dev = _dev()
    g = gold["planner"]
    s = g["state"]
    attrs = hr.activate(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"],
                        s["view_scores"], s["view_supports"], s["view_means"])
    r = O.GaussianRenderer(g["ext"].to(dev), g["K"].to(dev), tuple(a.to(dev) for a in attrs), torch.zeros(4, device=dev),
                           (0.001, 10.0), g["hw"], dev)
    out = r.render_view_all()
    depth, conf = out[1][:, 0].contiguous(), out[5][:, 0].contiguous()
    M = g["voxel_centers"].shape[0]
    gen = torch.Generator().manual_seed(2)
    for valid in (None, torch.rand(depth.shape, generator=gen) < 0.8):
        ex_ref, ei_ref = hr.view_utilities(depth.cpu(), conf.cpu(), g["voxel_centers"], g["unexplored"], g["ext"], g["K"],
                                           g["depth_range"], valid_mask=valid)
        ex, ei = ops.view_utility(depth, conf, g["voxel_centers"].to(dev), g["unexplored"].to(dev),
                                  torch.linalg.inv(g["ext"]).to(dev), g["K"].to(dev), g["depth_range"],
                                  valid=None if valid is None else valid.to(dev))
        assert float((ex.cpu() - ex_ref).abs().max()) <= 2.5 / M         # a borderline voxel or two may flip (fp32 order)
        assert torch.allclose(ei.cpu(), ei_ref, rtol=1e-5, atol=1e-7)
    # and against the reference planner's own numbers (our render vs the oracle's: 1e-4 rel)
    ex, ei = ops.view_utility(depth, conf, g["voxel_centers"].to(dev), g["unexplored"].to(dev),
                              torch.linalg.inv(g["ext"]).to(dev), g["K"].to(dev), g["depth_range"])
    assert float((ex.cpu() - g["utility_exploration"]).abs().max()) <= 4.5 / M
    assert torch.allclose(g["explore_weight"] * ex.cpu() + ei.cpu(), g["utility_confidence"], atol=2e-2, rtol=1e-3)


def test_planner_mirror_matches_reference_planner_outputs(gold):
    """active_gs_b200.planning.Confidence / Exploration.cal_utility against the numbers the reference's
    own planning/confidence.py and planning/exploration.py produced on the same map (fixture)."""
    from types import SimpleNamespace as ns
    from active_gs_b200 import planning
    from active_gs_b200.config import default_gaussian_map_config
    from active_gs_b200.gaussian_map import GaussianMap
    dev = _dev()
    g = gold["planner"]
    gm = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in g["state"].items():
        setattr(gm, k if k.startswith("view_") else "_" + k, v.clone().to(dev))
    vm = ns(voxel_centers=g["voxel_centers"], unexplored_mask=g["unexplored"])
    h, w = g["hw"]
    sim = ns(resolution=np.array([4 * h, 4 * w]), depth_range=g["depth_range"], intrinsic=g["K"][0], has_missing_surface=False)
    cfg = ns(render_ratio=0.25, explore_weight=g["explore_weight"])
    M = g["voxel_centers"].shape[0]
    u_c, _ = planning.Confidence(cfg, dev).cal_utility(gm, vm, g["ext"], sim)
    u_e, _ = planning.Exploration(cfg, dev).cal_utility(gm, vm, g["ext"], sim)
    assert float((u_e - g["utility_exploration"]).abs().max()) <= 4.5 / M
    assert torch.allclose(u_c, g["utility_confidence"], atol=2e-2, rtol=1e-3)
    assert int(torch.argmax(u_c)) == int(torch.argmax(g["utility_confidence"]))


def test_voxel_roi_matches_oracle_and_reference_fixture(gold):
    """ags_voxel_roi / planning.low_confidence_voxels vs VoxelMap.update_utility's voxel_normal (fixture)"""
    from types import SimpleNamespace as ns
    from active_gs_b200 import planning
    from active_gs_b200.config import default_gaussian_map_config
    from active_gs_b200.gaussian_map import GaussianMap
    from oracle import host_ref as hr
    dev = _dev()
    g = gold["voxel_roi"]
    gm = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in g["state"].items():
        setattr(gm, k if k.startswith("view_") else "_" + k, v.clone().to(dev))
    vm = ns(bbox=g["bbox"], size=g["size"], dim=g["dim"], min_gaussian_per_voxel=g["min_gaussian_per_voxel"])
    normal, mask = planning.low_confidence_voxels(vm, gm, g["confidence_thres"])
    count_ref, vn_ref, upd_ref = hr.low_confidence_voxels(g["state"], g["bbox"][0], g["size"], g["dim"],
                                                          g["min_gaussian_per_voxel"], g["confidence_thres"])
    assert torch.equal(mask.cpu(), upd_ref)
    assert torch.allclose(normal.cpu(), vn_ref, atol=1e-5)
    assert torch.allclose(normal.cpu(), g["voxel_normal"], atol=1e-5)
    from active_gs_b200 import ops
    count, _, _ = ops.voxel_roi(gm._means, gm._rotations, gm._opacities, gm.get_confidences, g["bbox"][0].tolist(),
                                g["size"].tolist(), g["dim"].tolist(), min_gaussian_per_voxel=g["min_gaussian_per_voxel"])
    assert torch.equal(count.cpu().long(), count_ref)
    # the whole update_utility (mapping/voxel_map.py:62-116): frontier | low-confidence voxels, next to free space
    from scipy.ndimage import binary_dilation, generate_binary_structure
    dim = [int(d) for d in g["dim"]]
    M = dim[0] * dim[1] * dim[2]
    gen = torch.Generator().manual_seed(9)
    vm.frontier_mask = torch.rand(M, generator=gen) < 0.05
    vm.free_mask = torch.rand(M, generator=gen) < 0.3
    for use_conf in (True, False):
        roi = planning.update_utility(vm, gm, use_conf, g["confidence_thres"])
        raw = vm.frontier_mask | (upd_ref if use_conf else torch.zeros(M, dtype=torch.bool))
        dil = binary_dilation(vm.free_mask.view(*dim).numpy(), structure=generate_binary_structure(3, 1))
        assert torch.equal(roi.cpu(), raw & torch.from_numpy(dil).view(-1))
        assert torch.allclose(vm.voxel_normal.cpu(), vn_ref if use_conf else torch.zeros(M, 3), atol=1e-5)


def test_map_store_external_tensors_public_prune_and_growth(tmp_path):
    """GaussianMap keeps the reference's attribute surface on top of the capacity buffers: tensors
    assigned from outside (load(), tests) are adopted, prune(mask) ORs the opacity test into the caller's
    mask in place (quirk Q5, mapping/gaussian_map.py:234-246), add_gaussians appends without touching the
    existing rows, and old references are not silently resized."""
    from active_gs_b200.config import default_gaussian_map_config
    from active_gs_b200.gaussian_map import GaussianMap
    dev = _dev()
    state = syn.make_room_scene(3000, seed=31)
    gm = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in state.items():
        setattr(gm, k if k.startswith("view_") else "_" + k, v.clone().to(dev))
    gm.save(str(tmp_path), "t")
    gm2 = GaussianMap(default_gaussian_map_config(), dev)
    gm2.load(str(tmp_path / "map_t.th"))
    # public prune with a caller-owned mask
    g = torch.Generator().manual_seed(1)
    mask = (torch.rand(3000, generator=g) < 0.3).to(dev)
    want_drop = mask.cpu() | (torch.sigmoid(state["opacities"]) < 0.1)
    before = {k: getattr(gm2, k if k.startswith("view_") else "_" + k).clone() for k in state}
    gm2.prune(mask)
    assert torch.equal(mask.cpu(), want_drop)
    n = int((~want_drop).sum())
    for k in state:
        t = getattr(gm2, k if k.startswith("view_") else "_" + k)
        assert t.shape[0] == n and torch.equal(t.cpu(), before[k].cpu()[~want_drop]), k
    assert gm2._harmonics.shape == (n, 1, 3)
    # spawn appends after the surviving rows
    f = _plane_frame(48, 64, 9)
    old_means = gm2._means
    kept = gm2._means.clone()
    torch.manual_seed(3)
    gm2.add_gaussians(f)
    n2 = gm2._means.shape[0]
    assert n2 > n and old_means.shape[0] == n
    assert torch.equal(gm2._means[:n], kept)
    assert torch.equal(gm2._scales[n:].cpu(), torch.tensor([0.0, 0.0, -1e10]).expand(n2 - n, 3))
    assert len(gm2.training_data) == 1 and gm2.training_performance.shape[0] == 1
    # externally replaced tensors (different length) are adopted by the next call
    for k, v in state.items():
        setattr(gm2, k if k.startswith("view_") else "_" + k, v[:500].clone().to(dev))
    gm2.add_gaussians(_plane_frame(48, 64, 10))
    assert gm2._means.shape[0] > 500 and torch.equal(gm2._means[:500].cpu(), state["means"][:500])
    assert gm2.get_attr()[0].shape[0] == gm2.view_scores.shape[0] == gm2._means.shape[0]
