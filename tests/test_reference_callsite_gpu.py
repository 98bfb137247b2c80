"""GPU: the reference's OWN, UNMODIFIED call site -- utils/operations.py:645-720 (render_cuda_core) and
:723-904 (GaussianRenderer) -- executed over the drop-in `diff_gaussian_rasterization_2d` module, forward and
autograd backward, compared with the oracle.

The reference file is located at run time (never copied into the repository): /root/reference in the build
container, or $AGS_REFERENCE_DIR / tests/_ref_tmp (an untracked, git-ignored directory that a gpurun call may
ship as test input and that is deleted afterwards).  Without it the test skips: the GPU box has no
/root/reference.  Absent third-party imports of that file (open3d, trimesh) are satisfied by empty stub
modules; the functions exercised here do not touch them.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

from oracle import host_ref as hr, rasterizer_ref as rr

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
NAMES = ["means", "scales", "rotations", "opacities", "harmonics"]


def _reference_operations():
    for base in [os.environ.get("AGS_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "tests", "_ref_tmp")]:
        if base and os.path.exists(os.path.join(base, "utils", "operations.py")):
            path = os.path.join(base, "utils", "operations.py")
            break
    else:
        pytest.skip("the reference's utils/operations.py is not available on this machine")
    for n in ["trimesh", "open3d"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    import diff_gaussian_rasterization_2d as drop_in
    assert drop_in.__file__.startswith(ROOT), "the drop-in module must be the repository's"
    spec = importlib.util.spec_from_file_location("ags_reference_operations", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.GaussianRasterizer is drop_in.GaussianRasterizer
    return mod, path


def test_reference_render_cuda_core_runs_unmodified_over_the_dropin():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    dev = torch.device("cuda:0")
    ref_ops, path = _reference_operations()
    print("  reference call site:", path)
    from active_gs_b200 import synthetic as syn
    state, ext, K = syn.make_c1_scene()
    H = W = 64
    leaves = [state[k].clone().to(dev).requires_grad_(True) for k in NAMES]
    attrs = hr.activate(*leaves[:4], leaves[4], state["view_scores"].to(dev), state["view_supports"].to(dev),
                        state["view_means"].to(dev))
    # the reference's renderer, exactly as mapping/gaussian_map.py:94-104 drives it
    out = ref_ops.GaussianRenderer(ext.to(dev), K.to(dev), attrs, torch.zeros(4, device=dev), (0.001, 10.0), (H, W),
                                   dev).render_view_all(require_grad=True)
    gt_rgb, gt_d = torch.rand(1, 3, H, W), 1.5 + torch.rand(1, 1, H, W)
    loss, _ = hr.train_loss(out[0], out[1], out[2], out[3], out[4], gt_rgb.to(dev), gt_d.to(dev))
    loss.backward()
    # oracle: the restated host half over the CPU rasterizer restatement
    cl = [state[k].clone().requires_grad_(True) for k in NAMES]
    attrs_c = hr.activate(*cl[:4], cl[4], state["view_scores"], state["view_supports"], state["view_means"])
    out_c = hr.render_view_all(rr.rasterize, ext, K, attrs_c, torch.zeros(4), (0.001, 10.0), (H, W), require_grad=True)
    loss_c, _ = hr.train_loss(out_c[0], out_c[1], out_c[2], out_c[3], out_c[4], gt_rgb, gt_d)
    loss_c.backward()
    for name, a, b in zip(["rgb", "depth", "normal", "opacity", "d2n", "confidence"], out[:6], out_c[:6]):
        a, b = a.detach().cpu(), b.detach()
        frac = ((a - b).abs() > 1e-4 * b.abs().max().clamp_min(1e-12)).float().mean().item()
        print(f"  {name:10s} max abs diff {float((a - b).abs().max()):.2e}  frac beyond 1e-4: {frac:.2e}")
        assert frac < 2e-3, name                      # threshold flips at the d2n / opacity masks only
    assert abs(float(loss.detach()) - float(loss_c.detach())) <= 1e-4 * abs(float(loss_c.detach()))
    for k, a, b in zip(NAMES, leaves, cl):
        l2 = (a.grad.cpu() - b.grad).norm() / b.grad.norm().clamp_min(1e-30)
        print(f"  grad {k:10s} l2_rel={float(l2):.3e}")
        assert l2 < 2e-4, k
    # forward-only call with the importance / front-only switches (post_processing, gaussian_map.py:183-192)
    with torch.no_grad():
        o2 = ref_ops.GaussianRenderer(ext.to(dev), K.to(dev), [t.detach() for t in attrs], torch.zeros(4, device=dev),
                                      (0.001, 10.0), (H, W), dev).render_view_all(require_importance=True, front_only=True)
    assert o2[7].dtype == torch.int32 and int(o2[7].sum()) > 0          # counts
