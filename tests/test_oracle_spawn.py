"""CPU: the oracle's restatement of the spawn step (mapping/gaussian_map.py:294-468), of the voxel
filter's guarantees and of the planner utilities is pinned to fixtures produced by EXECUTING the
reference's own Python (tests/golden/make_golden_spawn.py -> spawn_golden.pt)."""
import os
import pytest
import torch

from oracle import host_ref as hr, rasterizer_ref as rr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "spawn_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _check_new(c, new):
    sel = c["select"]
    assert int(sel.sum()) == new["means"].shape[0]
    assert torch.allclose(c["means"][sel], new["means"], atol=1e-6)
    assert torch.allclose(c["rotations"][sel], new["rotations"], atol=1e-5)
    assert torch.allclose(c["colors"][sel], new["harmonics"][:, 0, :], atol=0)
    n = new["means"].shape[0]
    assert torch.equal(new["scales"], torch.tensor([0.0, 0.0, -1e10]).expand(n, 3))
    assert float(new["opacities"].abs().max()) == 0.0 and float(new["view_supports"].abs().max()) == 0.0


def test_smooth_depth_matches_reference(gold):
    g = gold["smooth_depth"]
    assert torch.equal(hr.smooth_depth(g["depth"])[0], g["out"])


def test_spawn_first_keyframe(gold):
    g = gold["first"]
    c = hr.spawn_candidates(g["frame"])
    assert 0 < int(c["select"].sum()) < c["select"].numel()          # invalid / back-facing pixels dropped
    _check_new(c, g["new"])


def test_spawn_on_initialised_map(gold):
    g = gold["second"]
    s, f = g["state"], g["frame"]
    attrs = hr.activate(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"],
                        s["view_scores"], s["view_supports"], s["view_means"])
    with torch.no_grad():
        out = hr.render_view_all(rr.rasterize, f["extrinsic"][None], f["intrinsic"][None], attrs, torch.zeros(4),
                                 (0.001, 10.0), g["hw"])
    pred = dict(rgb=out[0][0], depth=out[1][0, 0], opacity=out[3][0, 0])
    c = hr.spawn_candidates(f, pred)
    first = hr.spawn_candidates(f)
    assert int(c["select"].sum()) < int(first["select"].sum())       # the map already explains part of the frame
    _check_new(c, g["new"])


def test_voxel_filter_properties(gold):
    for key in ["voxel", "voxel_coarse"]:
        g = gold[key]
        assert hr.voxel_filter_is_valid(g["points"], g["selected"], g["voxel_size"])
    g = gold["voxel_coarse"]
    assert g["selected"].numel() < g["points"].shape[0]              # the coarse case really merges points
    bad = g["selected"].clone()
    bad[0] = bad[1]
    assert not hr.voxel_filter_is_valid(g["points"], bad, g["voxel_size"])


def test_planner_utilities(gold):
    g = gold["planner"]
    s = g["state"]
    attrs = hr.activate(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"],
                        s["view_scores"], s["view_supports"], s["view_means"])
    with torch.no_grad():
        out = hr.render_view_all(rr.rasterize, g["ext"], g["K"], attrs, torch.zeros(4), (0.001, 10.0), g["hw"])
    depth, conf = out[1][:, 0], out[5][:, 0]
    for i in range(g["ext"].shape[0]):
        dv = depth[i].clone()
        dv[dv < 0.001] = 10000.0
        dv = dv.clamp(g["depth_range"][0], g["depth_range"][1])
        assert torch.equal(hr.voxel_visible_mask(g["voxel_centers"], g["ext"][i], g["K"][i], dv), g["visible"][i])
    explore, exploit = hr.view_utilities(depth, conf, g["voxel_centers"], g["unexplored"], g["ext"], g["K"],
                                         g["depth_range"])
    assert torch.allclose(explore, g["utility_exploration"], atol=1e-7)
    assert torch.allclose(g["explore_weight"] * explore + exploit, g["utility_confidence"], atol=1e-6)
    assert float(exploit.min()) > 0


def test_low_confidence_voxels(gold):
    g = gold["voxel_roi"]
    count, vn, upd = hr.low_confidence_voxels(g["state"], g["bbox"][0], g["size"], g["dim"], g["min_gaussian_per_voxel"],
                                              g["confidence_thres"])
    assert int(upd.sum()) > 20
    assert torch.equal(upd, g["voxel_normal"].norm(dim=1) > 0)
    assert torch.allclose(vn, g["voxel_normal"], atol=1e-6)
