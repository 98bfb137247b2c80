"""GPU parity of the fused image-loss kernel (K8) and the fused Adam (K7) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import host_ref as hr

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30), \
        (a - b).norm().item() / max(b.norm().item(), 1e-30)


def make_inputs(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g)
    rgb_gt = torch.rand(B, 3, H, W, generator=g)
    # piecewise-smooth depth so the TV depth gate (|dd|^2 <= 1e-4) opens on part of the image
    xs = torch.arange(W).float()[None, None, None, :] / W
    ys = torch.arange(H).float()[None, None, :, None] / H
    depth = 2.0 + 0.3 * xs + 0.2 * ys + 0.002 * torch.randn(B, 1, H, W, generator=g)
    depth[:, :, :, W // 2:] += 0.5
    depth_gt = depth + 0.05 * torch.randn(B, 1, H, W, generator=g)
    depth_gt[torch.rand(B, 1, H, W, generator=g) < 0.1] = -1.0
    normal = torch.randn(B, 3, H, W, generator=g) * 0.5 + torch.tensor([0.0, 0.0, -1.0])[None, :, None, None]
    opacity = torch.rand(B, 1, H, W, generator=g)
    opacity[torch.rand(B, 1, H, W, generator=g) < 0.15] = 0.0
    opacity[:, :, :2, :] = 0.005          # between the two thresholds (Q3)
    normal = normal * opacity              # rasterizer normals scale with coverage
    return rgb, normal, depth, opacity, rgb_gt, depth_gt


def oracle_loss(rgb, normal, depth, opacity, rgb_gt, depth_gt, fovs):
    """operations.py:714-718 post-processing + gaussian_map.py:106-124 via the restated host oracle.
    Evaluated in float64: where all four depth2normal difference vectors of a pixel vanish
    (isolated border pixel) normalize() divides by its eps=1e-12 and the fp32 autograd of the
    reference leaves 1e12-amplified terms that only cancel to ~1e-7 relative -- noise of order 1 in
    d_depth at those pixels.  The kernel returns the exact (cancelled) value, so is fp64 here."""
    B = rgb.shape[0]
    rgb, normal, depth, opacity, rgb_gt, depth_gt = [t.double() for t in
                                                     (rgb, normal, depth, opacity, rgb_gt, depth_gt)]
    leaves = [t.clone().requires_grad_(True) for t in (rgb, normal, depth)]
    r, n, d = leaves
    nus, d2ns = [], []
    for i in range(B):
        mask = opacity[i] > 1e-2
        nus.append(torch.nn.functional.normalize(n[i], dim=0) * mask)
        d2ns.append(hr.depth2normal(d[i], mask, fovs[i]))
    nu, d2n = torch.stack(nus), torch.stack(d2ns)
    total, perf = hr.train_loss(r, d, nu, opacity, d2n, rgb_gt, depth_gt)
    total.backward()
    return total.detach(), perf, nu.detach(), d2n.detach(), r.grad, n.grad, d.grad


@pytest.mark.parametrize("B,H,W", [(1, 24, 24), (3, 37, 53), (8, 48, 64)])
def test_fused_loss_matches_oracle(B, H, W):
    dev = _dev()
    from active_gs_b200 import ops
    ins = make_inputs(B, H, W, seed=B * 100 + H)
    fovs = torch.tensor([[1.0472, 0.8170]]).repeat(B, 1) + 0.01 * torch.arange(B)[:, None]
    total, perf, nu, d2n, g_rgb, g_n, g_d = oracle_loss(*ins, fovs)
    res = ops.loss_forward_backward(*[t.to(dev).contiguous() for t in ins], (0.5 * fovs).tan().to(dev))
    torch.cuda.synchronize()
    checks = [("total", res.total().reshape(1), total.reshape(1)), ("perf", res.frame_perf(), perf),
              ("normal_unit", res.normal_unit, nu), ("d2n", res.d2n, d2n), ("d_rgb", res.d_rgb, g_rgb),
              ("d_normal", res.d_normal, g_n), ("d_depth", res.d_depth, g_d)]
    ok = True
    for name, a, b in checks:
        emax, l2 = rel(a, b)
        print(f"  {name:12s} max_rel={emax:.3e} l2_rel={l2:.3e}")
        ok &= l2 < 1e-4 and emax < 2e-4
    assert ok


def test_fused_loss_vis_count_override_and_btotal():
    """multi-GPU form: frames split in two calls with the global visibility count and B_total must
    reproduce the single-call gradients (quirk Q1 couples frames through the mask sum)."""
    dev = _dev()
    from active_gs_b200 import ops
    B, H, W = 4, 20, 28
    ins = [t.to(dev).contiguous() for t in make_inputs(B, H, W, seed=5)]
    fovs = (0.5 * torch.tensor([[1.0, 0.8]])).tan().repeat(B, 1).to(dev)
    full = ops.loss_forward_backward(*ins, fovs)
    vis = (ins[3] > 1e-3).sum(0)[0].to(torch.int32).contiguous()
    parts = [ops.loss_forward_backward(*[t[s].contiguous() for t in ins], fovs[s].contiguous(),
                                       B_total=B, vis_count=vis) for s in (slice(0, 2), slice(2, 4))]
    for name in ["d_rgb", "d_normal", "d_depth"]:
        got = torch.cat([getattr(p, name) for p in parts])
        torch.testing.assert_close(got, getattr(full, name), rtol=1e-5, atol=1e-9)
    tot = sum(p.terms[:4] for p in parts)
    torch.testing.assert_close(tot, full.terms[:4], rtol=1e-5, atol=1e-8)


def test_fused_adam_matches_torch():
    dev = _dev()
    from active_gs_b200 import ops
    g = torch.Generator().manual_seed(0)
    N = 5003
    shapes = [(N, 3), (N, 3), (N, 4), (N,), (N, 1, 3)]
    lrs = [hr.LR["mean"], hr.LR["scale"], hr.LR["rotation"], hr.LR["opacity"], hr.LR["harmonic"]]
    p0 = [torch.randn(*s, generator=g) for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in p0]
    opt = hr.make_adam(*ref)
    ours = [p.clone().to(dev) for p in p0]
    m = [torch.zeros_like(p) for p in ours]
    v = [torch.zeros_like(p) for p in ours]
    step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    for it in range(1, 13):
        grads = [torch.randn(*s, generator=g) * (10.0 ** (-(it % 5))) for s in shapes]
        grads[1][:, 2] = 0.0                      # the inert third-scale lane
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        opt.step()
        gd = [gr.to(dev) for gr in grads]
        if it % 2:
            ops.adam_step(ours, gd, m, v, lrs, step=it)
            step_dev += 1
        else:
            ops.adam_step(ours, gd, m, v, lrs, step_dev=step_dev)   # device counter path
    torch.cuda.synchronize()
    assert int(step_dev.item()) == 12
    for a, b in zip(ours, ref):
        emax, l2 = rel(a, b)
        assert l2 < 1e-6 and emax < 1e-5, (emax, l2)
    # zero-gradient lane stays exactly inert (0/(0+1e-15) = 0)
    assert torch.equal(ours[1][:, 2].cpu(), p0[1][:, 2])


def test_device_bilateral_matches_cv2():
    """ags_smooth_depth restates cv2.bilateralFilter(depth, 15, 0.5, 20) as used by get_smooth_depth
    (utils/operations.py:161-169): compare with OpenCV itself (CPU) on depth images with holes."""
    dev = _dev()
    cv2 = pytest.importorskip("cv2")
    from active_gs_b200 import operations as O
    rng = np.random.default_rng(0)
    for (H, W) in [(96, 128), (37, 53), (480, 640)]:
        xs = np.arange(W)[None, :] / W
        ys = np.arange(H)[:, None] / H
        depth = (2 + 0.5 * xs + 0.3 * ys + 0.01 * rng.standard_normal((H, W))).astype(np.float32)
        depth[:, W // 2:] += 0.8
        depth[rng.random((H, W)) < 0.05] = -1.0
        depth[:4, :10] = -2.0
        ref = O.get_smooth_depth(depth)                      # the reference's CPU path (cv2)
        got = O.get_smooth_depth_device(torch.tensor(depth, device=dev)[None])[0].cpu().numpy()
        assert np.array_equal(got < 0, depth < 0) and np.all(got[depth < 0] == -1.0)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print(f"  bilateral {H}x{W}: max rel err vs cv2 {err:.2e}")
        assert err < 2e-5
    flat = torch.full((1, 20, 30), 1.5, device=dev)
    assert torch.equal(O.get_smooth_depth_device(flat), flat)   # max == min: copied unchanged


def test_fused_loss_frame_lists_padding_weights_and_no_maps():
    """(i) ground truth through per-frame pointers == stacked tensors (bitwise); (ii) a padded frame
    (weight 0) contributes nothing: the 3 real frames + 1 padded frame reproduce the 3-frame call and the
    padded frame's gradients are zero; (iii) want_maps=False gives the same gradients."""
    dev = _dev()
    from active_gs_b200 import ops
    B, H, W = 4, 37, 53
    ins = [t.to(dev).contiguous() for t in make_inputs(B, H, W, seed=11)]
    fovs = (0.5 * torch.tensor([[1.0, 0.8]])).tan().repeat(B, 1).to(dev)
    full = ops.loss_forward_backward(*ins, fovs)
    rgb_list = [ins[4][k].clone() for k in range(B)]
    d_list = [ins[5][k].clone() for k in range(B)]
    lst = ops.loss_forward_backward(*ins[:4], rgb_list, d_list, fovs, want_maps=False)
    assert lst.normal_unit is None and lst.d2n is None
    for name in ["d_rgb", "d_normal", "d_depth"]:
        assert torch.equal(getattr(lst, name), getattr(full, name)), name
    torch.testing.assert_close(lst.terms, full.terms, rtol=1e-5, atol=1e-9)       # block sums land in atomic order
    # second call through the cached argument struct with DIFFERENT frame tensors (pointers refreshed)
    perm = [2, 0, 3, 1]
    ins_p = [t[perm].contiguous() for t in ins]
    ref_p = ops.loss_forward_backward(*ins_p, fovs[perm].contiguous())
    lst = ops.loss_forward_backward(ins_p[0], ins_p[1], ins_p[2], ins_p[3], [rgb_list[k] for k in perm],
                                    [d_list[k] for k in perm], fovs, out=None, want_maps=False)
    torch.testing.assert_close(lst.d_depth, ref_p.d_depth, rtol=1e-6, atol=1e-10)
    # padding: frames 0..2 real, frame 3 padded
    three = ops.loss_forward_backward(*[t[:3].contiguous() for t in ins], fovs[:3].contiguous())
    w = torch.tensor([1.0, 1.0, 1.0, 0.0], device=dev)
    pad = ops.loss_forward_backward(*ins, fovs, B_total=3, frame_weight=w)
    for name in ["d_rgb", "d_normal", "d_depth"]:
        got = getattr(pad, name)
        torch.testing.assert_close(got[:3], getattr(three, name), rtol=1e-5, atol=1e-10)
        assert float(got[3].abs().max()) == 0.0
    torch.testing.assert_close(pad.terms[:4], three.terms[:4], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(pad.terms[4:10], three.terms[4:10], rtol=1e-5, atol=1e-9)


def test_adam_zero_grad_consumes_the_gradients():
    dev = _dev()
    from active_gs_b200 import ops
    g = torch.Generator().manual_seed(1)
    shapes = [(1001, 3), (1001, 3), (1001, 4), (1001,), (1001, 1, 3)]
    p = [torch.randn(*s, generator=g).to(dev) for s in shapes]
    q = [t.clone() for t in p]
    gr = [torch.randn(*s, generator=g).to(dev) for s in shapes]
    gq = [t.clone() for t in gr]
    m, v = [torch.zeros_like(t) for t in p], [torch.zeros_like(t) for t in p]
    m2, v2 = [torch.zeros_like(t) for t in p], [torch.zeros_like(t) for t in p]
    lrs = [1e-3] * 5
    ops.adam_step(p, gr, m, v, lrs, step=1, zero_grad=True)
    ops.adam_step(q, gq, m2, v2, lrs, step=1)
    for a, b, c in zip(p, q, gr):
        assert torch.equal(a, b) and float(c.abs().max()) == 0.0
    for c in gq:
        assert float(c.abs().max()) > 0.0
