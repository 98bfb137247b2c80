"""CPU: the host-half oracle (oracle/host_ref.py) against fixtures produced by executing the
reference's own Python (tests/golden/make_golden.py)."""
import math
import numpy as np
import torch
import pytest

from oracle import host_ref as hr, rasterizer_ref as rr
from active_gs_b200 import synthetic as syn


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_fov_and_projection_kat(golden):
    g = golden["cam"]
    fov = hr.get_fov(g["K"])
    close(fov, g["fov"])
    # known answers (SURVEY 8c KAT 1): 60 deg square, P[0,0]=1.7321, P[2,2]=1.0001
    assert abs(math.degrees(fov[0, 0]) - 60.0) < 1e-3 and abs(math.degrees(fov[0, 1]) - 60.0) < 1e-3
    P = hr.projection_matrix(torch.full((3,), 0.001), torch.full((3,), 10.0), fov[:, 0], fov[:, 1])
    close(P, g["P"])
    assert abs(P[0, 0, 0] - 1.7321) < 1e-3 and abs(P[0, 2, 2] - 1.0001) < 1e-4


def test_camera_setup_matches_reference_renderer(golden):
    g = golden["renderer"]
    fovs, view, proj, campos = hr.camera_setup(g["ext"], g["K"], (0.001, 10.0))
    close(fovs, g["fovs"]); close(view, g["view"]); close(proj, g["proj"]); close(campos, g["campos"])
    close(hr.raydir_map(g["K"][0], *g["hw"]), g["raydir"])
    # clip.w == z_view: a point 2 m in front of camera 0
    p_world = g["ext"][0] @ torch.tensor([0.0, 0.0, 2.0, 1.0])
    assert abs((p_world @ proj[0])[3] - 2.0) < 1e-5


def test_depth2normal(golden):
    for c in golden["depth2normal"]:
        out = hr.depth2normal(c["depth"], c["mask"], c["fov"])
        close(out, c["out"], rtol=1e-4, atol=1e-5)
    # KAT 2: flat fronto-parallel plane -> (0,0,-1)
    d = torch.full((1, 8, 8), 2.0)
    n = hr.depth2normal(d, torch.ones(1, 8, 8, dtype=torch.bool), (math.pi / 3, math.pi / 3))
    close(n[:, 4, 4], torch.tensor([0.0, 0.0, -1.0]))


def test_rotations(golden):
    g = golden["rot"]
    close(rr.quat_to_rotmat(g["q"]), g["R"])
    q = syn.normal2rotation(g["normals"])
    close(q, g["q_from_normal"], rtol=1e-4, atol=1e-5)
    # KAT 3: third column of R(q) is the normal (n = (0,0,-1) is the reference's degenerate
    # trace = -1 case, rotmat2quaternion operations.py:526-541 -- excluded)
    ok = torch.ones(64, dtype=torch.bool); ok[1] = False
    close(rr.quat_to_rotmat(q)[ok][:, :, 2], g["normals"][ok], rtol=1e-4, atol=1e-4)


def test_losses_and_q1(golden):
    g = golden["loss"]
    total, perf = hr.train_loss(g["rgb_p"], g["d_p"], g["n_p"], g["op"], g["d2n"], g["rgb_g"], g["d_g"])
    close(total, g["total"])
    close(hr.central_diff(g["n_p"]), g["central_diff_n"])
    close(hr.normal_tv_loss(g["n_p"], g["d_p"], g["d_g"] > 0), g["tv"])
    assert g["cons_shape"] == (3, 3, 10, 14)              # quirk Q1: (B,B,H,W)
    assert abs(hr.cal_psnr(g["rgb_p"], g["rgb_g"]) - g["psnr"]) < 1e-4


def test_activations(golden):
    g = golden["activate"]
    s = g["state"]
    attrs = hr.activate(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"],
                        s["view_scores"], s["view_supports"], s["view_means"])
    for a, b in zip(attrs, g["attrs"]):
        close(a, b)
    assert attrs[4][:, 2].abs().max() == 0               # flat disks: third scale exactly 0
    assert attrs[4][5, 0] == 0.05                        # clamp
    assert attrs[3][3] == min(1.0, max(0.0, float(s["view_scores"][3])))  # NaN norm -> 1 -> exp(0)


def test_sampler(golden):
    g = golden["sampler"]
    np.random.seed(g["seed"])
    for d in g["draws"]:
        ids = hr.WeightedSampler(12).next_ids(g["perf"])
        assert np.array_equal(ids, d)
        assert list(ids[:3]) == [9, 10, 11] and len(set(ids)) == 8   # KAT 6
    assert np.array_equal(hr.WeightedSampler(2).next_ids(g["perf"][:2]), g["draws_T2"])


def _run_train(g, steps, prune_interval):
    frames = g["frames"]
    state = {k: v.clone() for k, v in g["start"].items()}
    perf = g.get("perf0", torch.full((len(frames),), 10.0)).clone()
    np.random.seed(g["np_seed"])
    sampler = hr.WeightedSampler(len(frames))
    hw = tuple(frames[0]["rgb"].shape[1:])
    np.random.seed(g["np_seed"])
    perf_track = perf.clone()

    class _LazyBatches:
        """feeds ids drawn from the *current* tracked performance, like the reference loop"""
        def __iter__(self_inner):
            for _ in range(steps):
                yield sampler.next_ids(perf_track)

    names = ["means", "scales", "rotations", "opacities", "harmonics"]
    params = [torch.nn.Parameter(state[k].clone()) for k in names]
    opt = hr.make_adam(*params)
    for ids in _LazyBatches():
        rgb_gt = torch.stack([frames[i]["rgb"] for i in ids])
        d_gt = torch.stack([frames[i]["depth"] for i in ids])
        ext = torch.stack([frames[i]["extrinsic"] for i in ids])
        intr = torch.stack([frames[i]["intrinsic"] for i in ids])
        attrs = hr.activate(*params[:4], params[4], state["view_scores"], state["view_supports"],
                            state["view_means"])
        rgb, depth, normal, opacity, d2n, *_ = hr.render_view_all(
            rr.rasterize, ext, intr, attrs, torch.zeros(4), (0.001, 10.0), hw, require_grad=True)
        loss, pf = hr.train_loss(rgb, depth, normal, opacity, d2n, rgb_gt, d_gt)
        perf_track[ids] = pf
        loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
    for k, p in zip(names, params):
        state[k] = p.detach()
    hr.post_process(state, frames, torch.zeros(4), (0.001, 10.0), hw, prune_interval)
    return state, perf_track


def test_train_c1_matches_reference_gaussianmap(golden):
    """BASELINE config 1: 1k Gaussians, one 64x64 frame, 10 iterations, reference
    GaussianMap.train() (with the oracle as the native module) vs the restated loop."""
    g = golden["train_c1"]
    state, perf = _run_train(g, 10, 5)
    for k, v in g["end"].items():
        close(state[k], v, rtol=1e-4, atol=1e-5)
    close(perf, g["perf"], rtol=1e-4, atol=1e-6)


def test_train_multiframe_prune_matches_reference(golden):
    g = golden["train_multi"]
    state, perf = _run_train(g, 3, g["prune_interval"])
    assert state["means"].shape == g["end"]["means"].shape
    for k, v in g["end"].items():
        close(state[k], v, rtol=1e-4, atol=1e-5)
    close(perf, g["perf"], rtol=1e-4, atol=1e-6)


def test_adam_first_step_is_sign(golden):
    """KAT 7: eps=1e-15 -> first update = -lr*sign(g)."""
    p = torch.randn(100); gr = torch.randn(100) * 1e-3
    p2, m, v = hr.adam_step_ref(p, gr, torch.zeros(100), torch.zeros(100), 1e-2, 1)
    close(p2 - p, -1e-2 * torch.sign(gr), rtol=1e-5, atol=1e-8)
    q = torch.nn.Parameter(p.clone()); q.grad = gr.clone()
    torch.optim.Adam([q], lr=1e-2, eps=1e-15).step()
    close(q.detach(), p2, rtol=1e-6, atol=1e-7)


def test_dilate6_matches_scipy_binary_dilation():
    """planning.dilate6 == scipy.ndimage.binary_dilation with the 6-connected structure the reference uses
    (mapping/voxel_map.py:21,290-304)"""
    from scipy.ndimage import binary_dilation, generate_binary_structure
    from active_gs_b200.planning import dilate6
    g = torch.Generator().manual_seed(5)
    for shape in [(7, 5, 3), (1, 4, 6), (12, 12, 12)]:
        m = torch.rand(*shape, generator=g) < 0.15
        want = binary_dilation(m.numpy(), structure=generate_binary_structure(3, 1))
        assert np.array_equal(dilate6(m).numpy(), want)
