"""GPU parity: libags_b200.so (through the C ABI / drop-in module) vs the oracle.

Tolerance (BASELINE.json north_star): 1e-4 relative, fp32.  'Relative' is measured against the
tensor's scale (max |ref|): err = max|got-ref| / max|ref|, and as a 2-norm ratio.  The compositing
rule has hard thresholds (alpha<1/255, T<1e-4, power>0, integer radius) where two fp32 evaluation
orders can legitimately fall on different sides, so a small fraction of outlier elements is
tolerated for images; the norm-wise error must still pass.
"""
import math
import numpy as np
import pytest
import torch

from oracle import rasterizer_ref as rr, host_ref as hr
from active_gs_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


def report(name, got, ref, tol=TOL, outlier_frac=0.0):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    scale = max(ref.abs().max().item(), 1e-12)
    diff = (got - ref).abs()
    emax = diff.max().item() / scale if diff.numel() else 0.0
    outl = diff > tol * scale
    frac = outl.double().mean().item() if diff.numel() else 0.0
    # threshold flips (alpha<1/255, T<1e-4 ...) put a whole pixel / Gaussian on the other side of a
    # discontinuity: the norm-wise error is taken over the non-outlier elements, whose share is bounded
    keep = ~outl if outlier_frac > 0 else torch.ones_like(outl)
    l2 = ((got - ref) * keep).norm().item() / max(ref.norm().item(), 1e-12)
    ok = (l2 <= tol) and (frac <= outlier_frac if outlier_frac > 0 else emax <= tol)
    print(f"  {name:12s} max_rel={emax:.3e} l2_rel={l2:.3e} frac_out={frac:.2e} {'ok' if ok else 'FAIL'}")
    return ok


def setup_case(state, ext, K, hw, dev):
    attrs = hr.activate(state["means"], state["scales"], state["rotations"], state["opacities"],
                        state["harmonics"], state["view_scores"], state["view_supports"],
                        state["view_means"])
    fovs, view, proj, campos = hr.camera_setup(ext, K, (0.001, 10.0))
    return attrs, fovs, view, proj, campos


def run_both(attrs, fovs, view, proj, hw, dev, vi=0, bg=None, render_mask=None,
             require_importance=False, front_only=False, upstream_seed=0, check_grad=True,
             ref_dtype=torch.float32, only_ref=False):
    """Render view `vi` with the oracle (CPU, `ref_dtype`) and the drop-in module (GPU) and return
    outputs+grads."""
    from diff_gaussian_rasterization_2d import GaussianRasterizationSettings, GaussianRasterizer
    means, harm, opac, conf, scales, rots = attrs
    H, W = hw
    tan = (0.5 * fovs[vi]).tan()
    bg = torch.zeros(4) if bg is None else bg
    g = torch.Generator().manual_seed(upstream_seed)
    ups = [torch.randn(c, H, W, generator=g) for c in (3, 3, 1, 1, 1)]

    def leafs(device, dtype=torch.float32):
        return [t.detach().clone().to(device=device, dtype=dtype).requires_grad_(check_grad)
                for t in (means, opac[:, None], harm[:, 0, :], scales, rots)]

    # oracle
    m, o, c, s, r = leafs("cpu", ref_dtype)
    m2 = torch.zeros_like(m, requires_grad=check_grad)
    ups_ref = [u.to(ref_dtype) for u in ups]
    out_ref = rr.rasterize(m, m2, o, conf.to(ref_dtype), c, s, r, image_height=H, image_width=W,
                           tanfovx=float(tan[0]), tanfovy=float(tan[1]), bg=bg, viewmatrix=view[vi],
                           projmatrix=proj[vi], render_mask=render_mask, weight_thres=0.03,
                           require_importance=require_importance, front_only=front_only)
    grads_ref = None
    if check_grad:
        loss = sum((u * x).sum() for u, x in zip(ups_ref, out_ref[:5]))
        loss.backward()
        grads_ref = [m.grad, m2.grad, o.grad, c.grad, s.grad, r.grad]
    if only_ref:
        return out_ref, grads_ref
    # CUDA through the reference-facing module
    m, o, c, s, r = leafs(dev)
    m2 = torch.zeros_like(m, requires_grad=check_grad)
    settings = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=float(tan[0]), tanfovy=float(tan[1]), bg=bg.to(dev),
        scale_modifier=1.0, viewmatrix=view[vi].to(dev), projmatrix=proj[vi].to(dev), sh_degree=0,
        campos=torch.zeros(3, device=dev), prefiltered=False,
        render_mask=(torch.tensor([], device=dev) if render_mask is None else render_mask.to(dev)),
        weight_thres=0.03, debug=False,
        config=torch.tensor([1.0, 1.0, 1.0, float(require_importance), float(front_only)]).to(dev))
    out = GaussianRasterizer(settings)(means3D=m, means2D=m2, opacities=o, confidences=conf.to(dev),
                                       shs=None, colors_precomp=c, scales=s, rotations=r,
                                       cov3D_precomp=None)
    grads = None
    if check_grad:
        loss = sum((u.to(dev) * x).sum() for u, x in zip(ups, out[:5]))
        loss.backward()
        grads = [m.grad, m2.grad, o.grad, c.grad, s.grad, r.grad]
    return out_ref, grads_ref, out, grads


OUT_NAMES = ["rgb", "normal", "depth", "opacity", "confidence", "importance", "count", "radii"]
GRAD_NAMES = ["d_means3D", "d_means2D", "d_opacity", "d_colors", "d_scales", "d_rotations"]


def l2_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm().item() / max(b.norm().item(), 1e-300)


def compare_all(out_ref, grads_ref, out, grads, img_outliers=2e-3, grad_tol=TOL, grads_floor=None):
    """grads_floor: gradients of the SAME semantics evaluated in fp32 by the oracle; when given,
    `grads_ref` is the fp64 arbiter and a gradient passes if its error vs the arbiter is within
    max(tol, 2 x the fp32 oracle's own error) -- i.e. no worse than fp32 evaluation noise."""
    ok = True
    for k in range(5):
        ok &= report(OUT_NAMES[k], out[k], out_ref[k], outlier_frac=img_outliers)
    ok &= report("importance", out[5], out_ref[5], tol=1e-3, outlier_frac=5e-3)
    cnt_diff = (out[6].cpu() - out_ref[6]).abs()
    print(f"  count        max_abs_diff={int(cnt_diff.max()) if cnt_diff.numel() else 0} "
          f"frac_diff={(cnt_diff > 0).float().mean().item() if cnt_diff.numel() else 0:.2e}")
    ok &= (cnt_diff > 0).float().mean().item() <= 5e-3 if cnt_diff.numel() else True
    rad_diff = (out[7].cpu() - out_ref[7]).abs()
    print(f"  radii        frac_diff={(rad_diff > 0).float().mean().item() if rad_diff.numel() else 0:.2e}")
    ok &= (rad_diff > 0).float().mean().item() <= 1e-3 if rad_diff.numel() else True
    if grads is not None:
        for k, (n, a, b) in enumerate(zip(GRAD_NAMES, grads, grads_ref)):
            tol = grad_tol
            if grads_floor is not None:
                fl = l2_rel(grads_floor[k], b)
                tol = max(grad_tol, 2.0 * fl)
                print(f"  {n:12s} fp32-oracle-vs-fp64 l2_rel={fl:.3e} -> tol {tol:.2e}")
            ok &= report(n, a, b, tol=tol, outlier_frac=2e-3)
    return ok


def test_c1_forward_backward():
    """BASELINE config 1 scene: 1k Gaussians, 64x64 -- all 8 outputs and all 6 input gradients."""
    dev = _dev()
    state, ext, K = syn.make_c1_scene()
    attrs, fovs, view, proj, _ = setup_case(state, ext, K, (64, 64), dev)
    res = run_both(attrs, fovs, view, proj, (64, 64), dev, require_importance=True,
                   bg=torch.tensor([0.1, 0.2, 0.3, 0.0]))
    assert compare_all(*res)


def test_room_cut_20k_160x120():
    """20k-Gaussian 160x120 cut of the config-2 room (SURVEY 8d parity gate)."""
    dev = _dev()
    box = (6.0, 4.5, 2.7)
    state = syn.make_room_scene(20000, box=box, seed=1002)
    state["scales"][:, :2] += 1.2          # fewer surfels than C2 -> larger disks for the same coverage
    ext, K = syn.make_cameras(2, box=box, H=120, W=160, seed=2002)
    attrs, fovs, view, proj, _ = setup_case(state, ext, K, (120, 160), dev)
    for vi in range(2):
        # fp64 oracle = arbiter; fp32 oracle gives the evaluation-noise floor of the gradients
        out64, g64, out, g = run_both(attrs, fovs, view, proj, (120, 160), dev, vi=vi,
                                      ref_dtype=torch.float64)
        _, g32 = run_both(attrs, fovs, view, proj, (120, 160), dev, vi=vi, only_ref=True)
        assert compare_all(out64, g64, out, g, grads_floor=g32)


def test_ragged_image_and_masked_counts():
    """Image size not a multiple of the tile, render_mask + front_only + importance/count."""
    dev = _dev()
    box = (3.0, 2.5, 2.0)
    state = syn.make_room_scene(3000, box=box, seed=31, furniture=3)
    state["scales"][:, :2] += 1.5
    H, W = 37, 53
    ext, K = syn.make_cameras(1, box=box, H=H, W=W, hfov=75.0, seed=32)
    attrs, fovs, view, proj, _ = setup_case(state, ext, K, (H, W), dev)
    mask = (torch.rand(1, H, W, generator=torch.Generator().manual_seed(3)) > 0.3).float()
    res = run_both(attrs, fovs, view, proj, (H, W), dev, render_mask=mask, require_importance=True,
                   front_only=True)
    assert compare_all(*res)


def test_single_surfel_known_answer():
    """KAT 4: one opaque fronto-parallel surfel at z=2 -> centre depth 2, normal (0,0,-1),
    rgb = colour*w + bg*(1-w)."""
    dev = _dev()
    from diff_gaussian_rasterization_2d import GaussianRasterizationSettings, GaussianRasterizer
    H = W = 32
    K = syn.normalised_intrinsic(H, W, 60.0, 60.0)[None]
    fovs, view, proj, _ = hr.camera_setup(torch.eye(4)[None], K, (0.001, 10.0))
    tan = (0.5 * fovs[0]).tan()
    # pixel centres are at integer coordinates with centre (W-1)/2: put the surfel on pixel (16,16)
    fx = W / (2 * float(tan[0]))
    xoff = (16 - (W - 1) / 2) / fx * 2.0
    means = torch.tensor([[xoff, xoff, 2.0]], device=dev)
    col = torch.tensor([[0.9, 0.5, 0.1]], device=dev)
    bg = torch.tensor([0.2, 0.2, 0.2, 0.0], device=dev)
    opa = torch.tensor([[0.8]], device=dev)
    settings = GaussianRasterizationSettings(
        H, W, float(tan[0]), float(tan[1]), bg, 1.0, view[0].to(dev), proj[0].to(dev), 0,
        torch.zeros(3, device=dev), False, torch.tensor([], device=dev), 0.03, False,
        torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0], device=dev))
    rgb, normal, depth, opacity, conf, imp, cnt, radii = GaussianRasterizer(settings)(
        means3D=means, means2D=torch.zeros_like(means), opacities=opa,
        confidences=torch.tensor([0.7], device=dev), shs=None, colors_precomp=col,
        scales=torch.tensor([[0.05, 0.05, 0.0]], device=dev),
        rotations=torch.tensor([[1.0, 0.0, 0.0, 0.0]], device=dev), cov3D_precomp=None)
    w = 0.8
    assert abs(opacity[0, 16, 16].item() - w) < 1e-4
    assert abs(depth[0, 16, 16].item() - 2.0) < 1e-4
    torch.testing.assert_close(rgb[:, 16, 16].cpu(), torch.tensor([0.9, 0.5, 0.1]) * w + 0.2 * (1 - w),
                               rtol=1e-4, atol=1e-5)
    n = torch.nn.functional.normalize(normal[:, 16, 16], dim=0).cpu()
    torch.testing.assert_close(n, torch.tensor([0.0, 0.0, -1.0]), rtol=1e-4, atol=1e-5)
    assert abs(conf[0, 16, 16].item() - 0.7 * w) < 1e-4
    assert radii.item() > 0 and cnt.item() >= 1
    assert opacity[0, 0, 0].item() == 0.0 and depth[0, 0, 0].item() == 0.0
    assert abs(rgb[0, 0, 0].item() - 0.2) < 1e-6


def test_empty_and_overflow_paths():
    dev = _dev()
    from active_gs_b200.rasterizer import RenderBatch
    from active_gs_b200 import lib as L
    z = lambda *s: torch.zeros(*s, device=dev)
    eye = torch.eye(4, device=dev)[None]
    tan = torch.tensor([[0.577, 0.577]], device=dev)
    rb = RenderBatch(z(0, 3), z(0, 3), z(0, 4), z(0), z(0, 3), z(0), eye, eye, tan,
                     torch.tensor([0.3, 0.4, 0.5], device=dev), 20, 20).forward()
    assert torch.allclose(rb.rgb[0, :, 3, 3].cpu(), torch.tensor([0.3, 0.4, 0.5]))
    assert rb.opacity.abs().max().item() == 0
    # overflow: tiny capacity forces the retry path and must give identical images
    state, ext, K = syn.make_c1_scene()
    attrs, fovs, view, proj, _ = setup_case(state, ext, K, (64, 64), dev)
    means, harm, opac, conf, scales, rots = [t.to(dev) for t in attrs]
    tanf = (0.5 * fovs).tan().to(dev)
    args = (means, scales, rots, opac, harm[:, 0, :], conf, view.to(dev), proj.to(dev), tanf, z(4), 64, 64)
    big = RenderBatch(*args, inst_cap=1 << 20).forward()
    small = RenderBatch(*args, inst_cap=16)
    small.forward(check_overflow=False)
    torch.cuda.synchronize()
    st = small.stats.tolist()
    assert st[L.STAT_OVERFLOW] == 1 and st[L.STAT_INSTANCES] == big.stats.tolist()[L.STAT_INSTANCES]
    assert small.opacity.abs().max().item() == 0          # nothing rendered on overflow
    small.forward(check_overflow=True)
    assert small.inst_cap >= st[L.STAT_INSTANCES]
    assert torch.equal(small.rgb, big.rgb) and torch.equal(small.depth, big.depth)


def test_oversize_tile_merge_sort_path():
    """> 4096 instances in one tile exercises the chunk-sort + merge path of the tile sorter."""
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    N = 9000
    means = torch.stack([0.02 * torch.randn(N, generator=g), 0.02 * torch.randn(N, generator=g),
                         1.0 + 2.0 * torch.rand(N, generator=g)], 1)
    state = dict(means=means, scales=torch.tensor([[0.0, 0.0, -1e10]]).repeat(N, 1),
                 rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(N, 1) + 0.1 * torch.randn(N, 4, generator=g),
                 opacities=torch.full((N,), -4.0), harmonics=torch.rand(N, 1, 3, generator=g),
                 view_scores=torch.rand(N, generator=g), view_supports=torch.ones(N),
                 view_means=torch.zeros(N, 3))
    K = syn.normalised_intrinsic(16, 16, 60.0, 60.0)[None]
    attrs, fovs, view, proj, _ = setup_case(state, torch.eye(4)[None], K, (16, 16), dev)
    res = run_both(attrs, fovs, view, proj, (16, 16), dev)
    assert compare_all(*res)
