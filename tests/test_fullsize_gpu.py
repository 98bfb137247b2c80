"""GPU: BASELINE.json full-size configurations through size-independent properties (the oracle is
too slow there), plus the complete GaussianMap.update() loop (spawn -> train -> post-process/prune).

Properties checked at full size (config[1]: 200k Gaussians, 640x480, B=8; config[4]-like 1080p view):
  * forward determinism: two runs give bit-identical images and the same per-tile depth order
  * sortedness: every tile's instance list is ordered by (depth, id), all depths >= the near cull
  * physical ranges: opacity in [0,1], rgb within the convex hull of colours and bg, depth >= 0
  * opacity/T bookkeeping: opacity == 1 - final_T, n_contrib <= tile instance count
  * linearity of the backward in the upstream gradients: bwd(a*g1 + g2) == a*bwd(g1) + bwd(g2)
  * batching: rendering B views in one call == rendering them one by one (bitwise)
  * gradient of a view-0-only loss is unaffected by the other views of the batch
  * overflow: a too-small workspace is detected, nothing is rendered, the retry matches
"""
import numpy as np
import pytest
import torch

from active_gs_b200 import synthetic as syn
from active_gs_b200.config import default_gaussian_map_config

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


def scene(cfg_idx, B, dev, N=None):
    from active_gs_b200 import operations as O
    box, H, W, N0 = syn.ROOMS[cfg_idx]
    N = N or N0
    state = syn.make_room_scene(N, box=box, seed=1000 + cfg_idx)
    ext, K = syn.make_cameras(B, box=box, H=H, W=W, seed=2000 + cfg_idx)
    fovs, view, proj, campos, tanfov = O.camera_blocks(ext, K, (0.001, 10.0))
    return state, (view.to(dev), proj.to(dev), tanfov.to(dev)), (H, W)


def batch(state, cams, hw, dev, **kw):
    from active_gs_b200.rasterizer import RenderBatch
    from active_gs_b200 import lib as L
    N = state["means"].shape[0]
    s = {k: v.to(dev) for k, v in state.items()}
    conf = torch.rand(N, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    return RenderBatch(s["means"], s["scales"], s["rotations"], s["opacities"], s["harmonics"].reshape(N, 3), conf,
                       cams[0], cams[1], cams[2], torch.tensor([0.1, 0.2, 0.3, 0.0], device=dev), hw[0], hw[1],
                       param_mode=L.PARAMS_RAW, **kw)


def workspace_views(rb):
    """decode the instance tables of a RenderBatch workspace (layout of csrc/ags_common.cuh)"""
    N, B, H, W, cap = rb.N, rb.B, rb.H, rb.W, rb.inst_cap
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    al = lambda x: (x + 255) & ~255
    off = rb.ws_ptr - rb.workspace.data_ptr()
    sizes = [("geom0", B * N * 16), ("geom1", B * N * 16), ("feat0", B * N * 16), ("feat1", B * N * 16),
             ("rect", B * N * 8), ("vis_list", B * N * 4), ("dsplat", B * N * 64), ("tile_count", B * tiles * 4),
             ("tile_offset", B * tiles * 4), ("tile_fill", B * tiles * 4), ("counters", 32),
             ("inst_key", cap * 8), ("inst_key_alt", cap * 8), ("inst_sorted", cap * 4),
             ("final_T", B * H * W * 4), ("n_contrib", B * H * W * 4)]
    out = {}
    for name, nbytes in sizes:
        out[name] = rb.workspace[off:off + nbytes]
        off += al(nbytes)
    return out, tiles


def test_config2_forward_properties_and_sortedness():
    dev = _dev()
    state, cams, hw = scene(2, 8, dev)
    a = batch(state, cams, hw, dev).forward()
    b = batch(state, cams, hw, dev).forward()
    torch.cuda.synchronize()
    for n in ["rgb", "normal", "depth", "opacity", "confidence", "radii"]:
        assert torch.equal(getattr(a, n), getattr(b, n)), f"{n} not deterministic"
    assert float(a.opacity.min()) >= 0 and float(a.opacity.max()) <= 1.0 + 1e-6
    assert float(a.depth.min()) >= 0 and torch.isfinite(a.rgb).all() and torch.isfinite(a.normal).all()
    cmax = state["harmonics"].max().item()
    assert float(a.rgb.max()) <= max(cmax, 0.3) + 1e-4 and float(a.rgb.min()) >= -1e-6
    ws, tiles = workspace_views(a)
    B, N = a.B, a.N
    cnt = ws["tile_count"].view(torch.int32).cpu()
    offs = ws["tile_offset"].view(torch.int32).cpu()
    ids = ws["inst_sorted"].view(torch.int32)
    depth = ws["feat0"].view(torch.float32).view(B * N, 4)[:, 3]
    final_T = ws["final_T"].view(torch.float32).view(B, 1, *hw)
    assert torch.allclose(a.opacity, 1 - final_T, atol=1e-6)
    assert int(cnt.sum()) == int(a.stats[0]) and int(a.stats[1]) == 0
    ncon = ws["n_contrib"].view(torch.int32).view(B, *hw).cpu()
    g = torch.Generator().manual_seed(0)
    for t in torch.randint(0, B * tiles, (300,), generator=g).tolist() + [int(cnt.argmax())]:
        n, o, v = int(cnt[t]), int(offs[t]), t // tiles
        if n == 0:
            continue
        seg = ids[o:o + n].long()
        d = depth[v * N + seg].cpu()
        key = torch.stack([d, seg.cpu().float()], 1)
        assert float(d.min()) > 0.2
        assert bool(((d[1:] > d[:-1]) | ((d[1:] == d[:-1]) & (seg.cpu()[1:] > seg.cpu()[:-1]))).all()), f"tile {t} unsorted"
        ty, tx = divmod(t % tiles, (hw[1] + 15) // 16)
        assert int(ncon[v, ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16].max()) <= n


def test_config2_backward_linearity_and_batching():
    dev = _dev()
    state, cams, hw = scene(2, 4, dev, N=60000)
    rb = batch(state, cams, hw, dev).forward()
    g = torch.Generator(device=dev).manual_seed(3)
    mk = lambda c: torch.randn(4, c, *hw, device=dev, generator=g)
    g1 = [mk(3), mk(3), mk(1), mk(1), mk(1)]
    g2 = [mk(3), mk(3), mk(1), mk(1), mk(1)]
    alpha = 0.37
    o1 = rb.backward(*g1)
    o2 = rb.backward(*g2)
    o12 = rb.backward(*[alpha * x + y for x, y in zip(g1, g2)])
    for a1, a2, a12 in zip(o1[:5], o2[:5], o12[:5]):
        ref = alpha * a1 + a2
        err = (a12 - ref).norm() / ref.norm().clamp_min(1e-30)
        assert float(err) < 2e-5, float(err)
    # batching: view k alone == view k inside the batch (bitwise), for images and for gradients
    k = 2
    one = batch(state, tuple(c[k:k + 1] for c in cams), hw, dev).forward()
    for n in ["rgb", "normal", "depth", "opacity", "confidence"]:
        assert torch.equal(getattr(one, n)[0], getattr(rb, n)[k])
    z = lambda t: torch.zeros_like(t)
    only_k = [torch.cat([z(x[:k]), x[k:k + 1], z(x[k + 1:])]) for x in g1]
    gb = rb.backward(*only_k)
    g1v = one.backward(*[x[k:k + 1] for x in g1])
    for a_, b_ in zip(gb[:5], g1v[:5]):
        err = (a_ - b_).norm() / b_.norm().clamp_min(1e-30)
        assert float(err) < 1e-5


def test_1080p_view_overflow_retry_and_counts():
    """config[4]-shaped single view (1920x1080, 300k surfels): overflow detection + retry, and the
    importance/count invariants under a render mask."""
    dev = _dev()
    from active_gs_b200 import lib as L
    state, cams, hw = scene(5, 1, dev, N=300000)
    mask = (torch.rand(1, *hw, device=dev) > 0.5).float()
    small = batch(state, cams, hw, dev, inst_cap=1000, render_mask=mask, require_importance=True)
    small.forward(check_overflow=False)
    torch.cuda.synchronize()
    st = small.stats.tolist()
    assert st[L.STAT_OVERFLOW] == 1 and st[L.STAT_INSTANCES] > 1000
    assert float(small.opacity.abs().max()) == 0
    small.forward(check_overflow=True)
    full = batch(state, cams, hw, dev, render_mask=mask, require_importance=True).forward()
    assert torch.equal(small.rgb, full.rgb) and torch.equal(small.count, full.count)
    cnt = full.count[0]
    assert int(cnt.max()) <= int(mask.sum()) and int(cnt.min()) >= 0
    assert bool((cnt[full.radii[0] == 0] == 0).all())
    assert bool(((full.importance[0] > 0) == (cnt > 0)).all())
    nomask = batch(state, cams, hw, dev, require_importance=True).forward()
    assert bool((nomask.count >= full.count).all())


def test_update_loop_spawn_train_prune():
    """GaussianMap.update() on a stream of keyframes (mapping/gaussian_map.py:62-64): spawn from the
    RGB-D frame, 10 optimisation steps, confidence bookkeeping, prune on the 5th keyframe.  PSNR of
    the re-rendered keyframes must improve while training and stay high after pruning."""
    dev = _dev()
    from active_gs_b200 import operations as O
    from active_gs_b200.gaussian_map import GaussianMap
    box, H, W = (6.0, 4.5, 2.7), 120, 160
    gen = syn.make_room_scene(40000, box=box, seed=5)
    gen["scales"][:, :2] += 0.9
    ext, K = syn.make_cameras(5, box=box, H=H, W=W, seed=6)
    src = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in gen.items():
        setattr(src, k if k.startswith("view_") else "_" + k, v.to(dev))
    gm = GaussianMap(default_gaussian_map_config(), dev)
    np.random.seed(0); torch.manual_seed(0)
    sizes, psnrs = [], []
    for i in range(5):
        with torch.no_grad():
            out = O.GaussianRenderer(ext[i:i + 1].to(dev), K[i:i + 1].to(dev), src.get_attr(), src.background_color,
                                     (0.001, 10.0), (H, W), dev).render_view_all()
        depth = torch.where(out[3][0] > 0.5, out[1][0], torch.full_like(out[1][0], -1.0))
        frame = dict(rgb=out[0][0].clamp(0, 1), depth=depth, extrinsic=ext[i].to(dev), intrinsic=K[i].to(dev),
                     depth_range=torch.tensor([0.0, 5.0]))
        gm.update(frame)
        sizes.append(gm._means.shape[0])
        with torch.no_grad():
            rr_ = O.GaussianRenderer(ext[:i + 1].to(dev), K[:i + 1].to(dev), gm.get_attr(), gm.background_color,
                                     (0.001, 10.0), (H, W), dev).render_view_all()
        gt = torch.stack([f["rgb"] for f in gm.training_data])
        psnrs.append(-10 * torch.log10(((rr_[0] - gt) ** 2).mean() + 1e-8).item())
        assert torch.isfinite(gm._means).all() and torch.isfinite(gm._opacities).all()
    print("  map sizes", sizes, "PSNR", ["%.2f" % p for p in psnrs])
    assert sizes[0] > 1000 and gm.is_init and len(gm.training_data) == 5
    assert gm.view_supports.max() >= 1 and gm.view_scores.max() > 0
    assert psnrs[-1] > 20.0 and min(psnrs) > 15.0
    l0, l1 = gm.last_train_log[0][0], gm.last_train_log[-1][0]
    assert l1 < l0


def test_planner_shaped_batch_100_views_128():
    """the planners render ~100 candidate views at 128x128 per step, forward only
    (planning/confidence.py:15-46).  One GaussianRenderer call = one launch chain for all views;
    every view must equal the same view rendered alone."""
    dev = _dev()
    from active_gs_b200 import operations as O
    from active_gs_b200.gaussian_map import GaussianMap
    box, H, W, N, V = (6.0, 4.5, 2.7), 128, 128, 100000, 100
    st = syn.make_room_scene(N, box=box, seed=21)
    gm = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in st.items():
        setattr(gm, k if k.startswith("view_") else "_" + k, v.to(dev))
    ext, K = syn.make_cameras(V, box=box, H=H, W=W, hfov=60.0, seed=22)
    with torch.no_grad():
        r = O.GaussianRenderer(ext.to(dev), K.to(dev), gm.get_attr(), gm.background_color, (0.001, 10.0), (H, W), dev)
        all_ = r.render_view_all()
        assert all_[0].shape == (V, 3, H, W) and all_[8].shape == (N,)
        assert torch.isfinite(all_[0]).all() and float(all_[3].max()) <= 1 + 1e-6
        for i in (0, 37, 99):
            one = r.render_view(i)
            for k in range(6):
                assert torch.equal(one[k], all_[k][i]), (i, k)
        # confidence map is a blend of per-Gaussian confidences in [0,1]
        assert float(all_[5].min()) >= 0 and float(all_[5].max()) <= 1 + 1e-5


def test_checkpoint_roundtrip(tmp_path):
    """mapping/gaussian_map.py:491-527: the .th dict keeps the reference's keys and raw tensors."""
    dev = _dev()
    from active_gs_b200.gaussian_map import GaussianMap
    st = syn.make_room_scene(5000, seed=3)
    gm = GaussianMap(default_gaussian_map_config(), dev)
    for k, v in st.items():
        setattr(gm, k if k.startswith("view_") else "_" + k, v.to(dev))
    gm.save(str(tmp_path), index=7)
    d = torch.load(tmp_path / "map_7.th", weights_only=False)
    assert set(d.keys()) == {"means", "scales", "harmonics", "opacities", "rotations", "view_scores", "view_supports",
                             "view_means", "near", "far", "use_view_direction", "background_color", "scale_factor"}
    g2 = GaussianMap(default_gaussian_map_config(), dev)
    g2.load(str(tmp_path / "map_7.th"))
    for a, b in zip(gm.get_params(), g2.get_params()):
        assert torch.equal(a, b)
    assert g2.is_init and g2.scene_far == 10.0
