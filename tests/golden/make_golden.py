"""Generate tests/golden/host_golden.pt by EXECUTING THE REFERENCE'S OWN PYTHON (run in the build
container only; /root/reference does not exist on the GPU box, so the output file is committed).

    python tests/golden/make_golden.py

What is pinned:  every host-side function of the hot path that lives in /root/reference --
get_fov, get_projection_matrix, GaussianRenderer.__init__ camera products, depth2normal (incl.
quirk Q2), quaternion_to_matrix, normal2rotation, the three losses, get_attr activations, the
WeightedSampler draw, and GaussianMap.train() end-to-end (loss composition incl. quirk Q1,
track_performance, Adam groups, post_processing bookkeeping).

What cannot be pinned: the native rasterizer itself (absent, un-pinned dependency).  Where the
reference calls `diff_gaussian_rasterization_2d`, the module is satisfied by
oracle/rasterizer_ref.py -- so the GaussianMap.train() fixture pins everything AROUND the native
call given our specification of it.

Absent third-party imports of the reference (open3d, trimesh, torchmetrics, imgviz, ...) are
replaced by empty stub modules; none of them is touched by the functions exercised here.
"""
import os
import sys
import types
import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import rasterizer_ref as rr  # noqa: E402
from active_gs_b200 import synthetic as syn  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def to(self, *a, **k):
        return self

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party object was called")


def import_reference():
    for n in ["trimesh", "open3d", "imgviz"]:
        _stub(n)
    _stub("torchmetrics", StructuralSimilarityIndexMeasure=_Anything)
    _stub("torchmetrics.image")
    _stub("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Anything)
    _stub("diff_gaussian_rasterization_2d",
          GaussianRasterizationSettings=rr.GaussianRasterizationSettings,
          GaussianRasterizer=rr.GaussianRasterizer)

    class TextColors:
        CYAN = RESET = ""

    _stub("utils.common", TextColors=TextColors, Camera=_Anything, Mapper2Gui=_Anything,
          FakeQueue=_Anything)
    import utils.operations as ops
    import mapping.utils as mutils
    import mapping.gaussian_map as gmap
    return ops, mutils, gmap


def cfg_namespace():
    """config/mapper/incremental.yaml:12-32 as attribute objects."""
    ns = types.SimpleNamespace
    return ns(bound=[0.001, 10.0], background=[0.0, 0.0, 0.0, 0.0], sparse_ratio=0.1,
              error_thres=0.25, scale_factor=0.01, optimization_steps=10, prune_interval=5,
              use_view_distribution=True,
              sampler=ns(sampler_type="weighted", batch_size=8, active_size=3),
              optimizer=ns(mean_lr=0.0005, rotation_lr=0.0005, opacity_lr=0.01, scale_lr=0.01,
                           harmonic_lr=0.0001))


def small_multiframe_case():
    """A 600-Gaussian 48x32 room with T=6 keyframes (exercises the sampler and Q1 with B>1 and a
    non-square image for Q2)."""
    box = (3.0, 2.5, 2.0)
    state = syn.make_room_scene(600, box=box, seed=77, furniture=2)
    # bigger disks so a 48x32 image is covered
    state["scales"][:, :2] += 2.0
    ext, K = syn.make_cameras(6, box=box, H=32, W=48, hfov=70.0, seed=78)
    return state, ext, K, (32, 48)


def make_frames(gmap, ops, state, ext, K, hw, seed):
    """Ground-truth dataframes rendered from the generating scene with the oracle."""
    gm = gmap.GaussianMap(cfg_namespace(), "cpu")
    load_state(gm, state)
    with torch.no_grad():
        rgb, depth, *_ = ops.GaussianRenderer(ext, K, gm.get_attr(), gm.background_color,
                                              (0.001, 10.0), hw, "cpu").render_view_all()
    frames = []
    for i in range(ext.shape[0]):
        d = syn.noisy_depth(depth[i], seed=seed + i)
        frames.append(dict(rgb=rgb[i].clamp(0, 1), depth=d, extrinsic=ext[i], intrinsic=K[i],
                           depth_range=torch.tensor([0.0, 5.0])))
    return frames


def load_state(gm, state):
    gm._means, gm._scales = state["means"].clone(), state["scales"].clone()
    gm._rotations, gm._opacities = state["rotations"].clone(), state["opacities"].clone()
    gm._harmonics = state["harmonics"].clone()
    gm.view_scores, gm.view_supports = state["view_scores"].clone(), state["view_supports"].clone()
    gm.view_means = state["view_means"].clone()


def dump_state(gm):
    return dict(means=gm._means.detach().clone(), scales=gm._scales.detach().clone(),
                rotations=gm._rotations.detach().clone(), opacities=gm._opacities.detach().clone(),
                harmonics=gm._harmonics.detach().clone(), view_scores=gm.view_scores.clone(),
                view_supports=gm.view_supports.clone(), view_means=gm.view_means.clone())


def main():
    ops, mutils, gmap = import_reference()
    torch.manual_seed(0)
    G = {}

    # ---- camera helpers
    Ks = torch.stack([syn.normalised_intrinsic(512, 512, 60.0, 60.0),
                      syn.normalised_intrinsic(480, 640, 60.0),
                      syn.normalised_intrinsic(720, 1280, 60.0)])
    fov = ops.get_fov(Ks)
    near, far = torch.full((3,), 0.001), torch.full((3,), 10.0)
    G["cam"] = dict(K=Ks, fov=fov, P=ops.get_projection_matrix(near, far, fov[:, 0], fov[:, 1]))
    ext, K = syn.make_cameras(4, H=480, W=640, seed=5)
    zero = torch.zeros(0, 3)
    r = ops.GaussianRenderer(ext, K, (zero,) * 6, torch.zeros(4), (0.001, 10.0), (30, 40), "cpu")
    G["renderer"] = dict(ext=ext, K=K, fovs=r.fovs, view=r.view_matrices, proj=r.projection_matrices,
                         campos=r.cam_pos, raydir=r.raydir_map, hw=(30, 40))

    # ---- depth2normal: flat, tilted, random; square and non-square (Q2)
    cases = []
    for (H, W, fv) in [(8, 8, (np.pi / 3, np.pi / 3)), (12, 20, (1.0472, 0.8170)), (16, 9, (0.7, 1.1))]:
        xs = torch.arange(W, dtype=torch.float32)[None, :].expand(H, W)
        for name, d in [("flat", torch.full((H, W), 2.0)), ("tilt", 2.0 + 0.05 * xs),
                        ("rand", 1.0 + torch.rand(H, W))]:
            d = d[None].clone()
            m = torch.rand(1, H, W) > 0.15
            cases.append(dict(name=name, depth=d, mask=m, fov=fv,
                              out=ops.depth2normal(d, m, fv)))
    G["depth2normal"] = cases

    # ---- rotations
    q = torch.nn.functional.normalize(torch.randn(64, 4), dim=-1)
    nrm = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
    nrm[0] = torch.tensor([1.0, 0.0, 0.0])
    nrm[1] = torch.tensor([0.0, 0.0, -1.0])
    q_n, R_n = ops.normal2rotation(nrm)
    G["rot"] = dict(q=q, R=ops.quaternion_to_matrix(q), normals=nrm, q_from_normal=q_n, R_from_normal=R_n)

    # ---- losses on random images
    B, H, W = 3, 10, 14
    rgb_p, rgb_g = torch.rand(B, 3, H, W), torch.rand(B, 3, H, W)
    d_p = 1 + torch.rand(B, 1, H, W)
    d_p[:, :, :, :7] = 2.0                                     # flat region so the depth-edge gate opens
    d_g = 1 + torch.rand(B, 1, H, W)
    d_g[d_g < 1.1] = -1.0
    n_p = torch.nn.functional.normalize(torch.randn(B, 3, H, W), dim=1)
    d2n = torch.nn.functional.normalize(torch.randn(B, 3, H, W), dim=1)
    op = torch.rand(B, 1, H, W)
    op[op < 0.2] = 0.0
    m_vis, m_d = op > 1e-3, d_g > 0.0
    l_rgb = mutils.l1_loss_fc_mask(rgb_p, rgb_g, m_vis)
    l_d = mutils.l1_loss_fc_mask(d_p, d_g, m_d)
    tv = mutils.normal_tv_loss_fc(n_p, d_p, m_d)
    cons = (mutils.cons_loss_fc(n_p, d2n) * m_vis.long())
    total = l_rgb.mean() + 0.8 * l_d.mean() + 0.1 * cons.mean() + 0.1 * tv     # gaussian_map.py:113-124
    G["loss"] = dict(rgb_p=rgb_p, rgb_g=rgb_g, d_p=d_p, d_g=d_g, n_p=n_p, d2n=d2n, op=op,
                     l_rgb_mean=l_rgb.mean(), l_d_mean=l_d.mean(), tv=tv, cons_shape=tuple(cons.shape),
                     cons_mean=cons.mean(), total=total, central_diff_n=mutils.central_diff(n_p),
                     psnr=mutils.cal_psnr(rgb_p, rgb_g))

    # ---- activations
    state, ext1, K1 = syn.make_c1_scene()
    state["view_means"][3] = float("nan")
    state["scales"][5, :2] = 3.0                                # hits the 0.05 clamp
    gm = gmap.GaussianMap(cfg_namespace(), "cpu")
    load_state(gm, state)
    G["activate"] = dict(state=state, attrs=[a.clone() for a in gm.get_attr()],
                         normals=gm.get_normals.clone())

    # ---- sampler
    np.random.seed(1234)
    perf = torch.tensor([10.0, 0.3, 0.7, 10.0, 0.2, 0.9, 1.5, 10.0, 0.4, 10.0, 10.0, 10.0])
    frames = [dict(rgb=torch.zeros(3, 2, 2), depth=torch.zeros(1, 2, 2), extrinsic=torch.eye(4),
                   intrinsic=torch.eye(3)) for _ in range(12)]
    cfgs = cfg_namespace().sampler
    draws = [mutils.WeightedSampler(cfgs, frames).next_frames(perf)[1] for _ in range(4)]
    draws2 = mutils.WeightedSampler(cfgs, frames[:2]).next_frames(perf[:2])[1]
    G["sampler"] = dict(seed=1234, perf=perf, draws=[np.asarray(d) for d in draws], draws_T2=np.asarray(draws2))

    # ---- GaussianMap.train() on BASELINE config 1 (1k Gaussians, one 64x64 frame, 10 iters)
    state, ext1, K1 = syn.make_c1_scene()
    frames = make_frames(gmap, ops, state, ext1, K1, (64, 64), seed=500)
    start = syn.perturb_state(state, seed=3001)
    gm = gmap.GaussianMap(cfg_namespace(), "cpu")
    load_state(gm, start)
    gm.training_data = frames
    gm.training_performance = torch.full((1,), 10.0)
    np.random.seed(1001)
    gm.train()
    G["train_c1"] = dict(start=start, frames=frames, end=dump_state(gm),
                         perf=gm.training_performance.clone(), np_seed=1001)

    # ---- GaussianMap.train(steps=3) on a multi-frame non-square case (sampler + Q1 + Q2 + prune)
    state, ext6, K6, hw = small_multiframe_case()
    frames = make_frames(gmap, ops, state, ext6, K6, hw, seed=600)
    start = syn.perturb_state(state, seed=3003)
    gm = gmap.GaussianMap(cfg_namespace(), "cpu")
    load_state(gm, start)
    gm.training_data = frames
    gm.training_performance = torch.tensor([0.8, 0.2, 10.0, 0.5, 10.0, 10.0])
    np.random.seed(1003)
    gm.prune_interval = 6                                     # T=6 -> post_processing renders all + prunes
    gm.train(steps=3)
    G["train_multi"] = dict(start=start, frames=frames, end=dump_state(gm), hw=hw,
                            perf=gm.training_performance.clone(), np_seed=1003,
                            perf0=torch.tensor([0.8, 0.2, 10.0, 0.5, 10.0, 10.0]), prune_interval=6)

    out = os.path.join(os.path.dirname(__file__), "host_golden.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
    print("train_c1 perf", G["train_c1"]["perf"], "N end", G["train_c1"]["end"]["means"].shape[0])
    print("train_multi perf", G["train_multi"]["perf"], "N end", G["train_multi"]["end"]["means"].shape[0])


if __name__ == "__main__":
    main()
