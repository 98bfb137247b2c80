"""BENCH-ONLY GPU-class baseline: the reference's own call chain on a straightforward 3DGS-lineage
rasterizer (baseline/gpu_naive.cu), so that the north_star target ">= 1.5x the reference CUDA rasterizer"
has a GPU denominator (the real extension, envs/requirements.txt:15, is not available).

What one iteration does -- the structure of /root/reference/mapping/gaussian_map.py:76-127 on a GPU:
    torch activations with autograd (get_attr, :529-581)
    -> B sequential single-view renders (utils/operations.py:853-892), each: per-Gaussian projection,
       device-wide scan, host read of the instance count, ONE global cub radix sort of (tile | depth) keys,
       one CTA per tile compositing every splat of the tile at every pixel
    -> torch post-processing (normalise, depth2normal) and the four loss terms as ~50 ATen kernels (:106-124)
    -> autograd backward: ATen backward kernels, per-pixel-atomic rasterizer backward, projection backward
    -> torch.optim.Adam, 5 parameter groups, eps 1e-15 (:259-292).
Nothing under active_gs_b200/ or diff_gaussian_rasterization_2d/ imports this module; bench.py (--impl
gpu_naive) and tests/test_gpu_naive_gpu.py are its only users.
"""
import ctypes as C
import os
import subprocess
import sys

import torch
import torch.nn.functional as F

from active_gs_b200 import lib as L
from active_gs_b200 import operations as O
from active_gs_b200.rasterizer import RenderBatch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gpu_naive.cu")
OUT = os.path.join(HERE, "libags_naive.so")
_lib = None


def build(force=False):
    deps = [SRC, os.path.join(HERE, "..", "active_gs_b200", "csrc", "ags_common.cuh"),
            os.path.join(HERE, "..", "include", "ags_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
    print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return OUT


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT):
            raise RuntimeError(f"{OUT} is missing: run __graft_entry__.build()")
        h = C.CDLL(OUT)
        h.naive_scratch_bytes.restype = C.c_size_t
        h.naive_scratch_bytes.argtypes = [C.c_int32] * 5
        h.naive_forward.restype = C.c_longlong
        h.naive_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32]
        h.naive_backward.restype = C.c_int
        h.naive_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        h.naive_launch_count.restype = C.c_ulonglong
        h.naive_last_error.restype = C.c_char_p
        _lib = h
    return _lib


_cap_hint = [6.0]        # instances per Gaussian that were enough last time


class NaiveView:
    """One view through the lineage pipeline.  The projection (K1) and its backward (K6) are the product's
    kernels, run stage by stage; binning / sort / compositing are baseline/gpu_naive.cu."""

    def __init__(self, means, scales, rots, opac, colors, conf, view, proj, tanfov, bg, H, W):
        self.rb = RenderBatch(means, scales, rots, opac, colors, conf, view.reshape(1, 16), proj.reshape(1, 16),
                              tanfov.reshape(1, 2), bg, H, W, inst_cap=4096, with_importance=False)
        self.N, self.H, self.W = self.rb.N, H, W
        self.cap = int(_cap_hint[0] * self.N) + 4096
        self._alloc()

    def _alloc(self):
        nb = lib().naive_scratch_bytes(self.N, 1, self.H, self.W, self.cap)
        self.nws = torch.empty(nb + 256, dtype=torch.uint8, device=self.rb.dev)
        self.nws_ptr, self.nws_bytes = (self.nws.data_ptr() + 255) & ~255, nb

    def forward(self):
        p, nl = L.load(), lib()
        a = self.rb._args()
        self.args = a
        L.check(p.ags_render_stage(C.byref(a), None, 0), "clear")
        L.check(p.ags_render_stage(C.byref(a), None, 1), "project_fwd")
        total = nl.naive_forward(C.byref(a), self.nws_ptr, self.nws_bytes, self.cap)
        if total > self.cap:
            self.cap = int(total * 1.25) + 4096
            self._alloc()
            total = nl.naive_forward(C.byref(a), self.nws_ptr, self.nws_bytes, self.cap)
        if total < 0:
            raise RuntimeError("naive_forward: " + nl.naive_last_error().decode())
        self.instances = int(total)
        _cap_hint[0] = max(_cap_hint[0] * 0.9, 1.5 * total / max(self.N, 1), 1.0)
        return self

    def backward(self, d_rgb, d_normal, d_depth, d_opacity, d_conf):
        p, nl = L.load(), lib()
        N, dev = self.N, self.rb.dev
        o = dict(device=dev, dtype=torch.float32)
        ups = [None if t is None else t.detach().float().contiguous() for t in (d_rgb, d_normal, d_depth, d_opacity, d_conf)]
        g = L.RenderGradArgs()
        g.d_rgb, g.d_normal, g.d_depth, g.d_opacity, g.d_confidence = [L.ptr(t) for t in ups]
        dm, ds, dr = torch.empty(N, 3, **o), torch.empty(N, 3, **o), torch.empty(N, 4, **o)
        do, dc = torch.empty(N, **o), torch.empty(N, 3, **o)
        g.d_means3D, g.d_scales, g.d_rotations = L.ptr(dm), L.ptr(ds), L.ptr(dr)
        g.d_opacities, g.d_colors, g.d_means2D = L.ptr(do), L.ptr(dc), None
        g.accumulate, g.clear_records = 0, 1
        a = self.args
        a.stream = L.current_stream(dev)
        rc = nl.naive_backward(C.byref(a), C.byref(g), self.nws_ptr, self.cap)
        if rc:
            raise RuntimeError("naive_backward: " + nl.naive_last_error().decode())
        L.check(p.ags_render_stage(C.byref(a), C.byref(g), 5), "project_bwd")
        return dm, ds, dr, do, dc


class _NaiveRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, colors, opac, scales, rots, conf, cam):
        view, proj, tanfov, bg, H, W = cam
        nv = NaiveView(means, scales, rots, opac, colors, conf, view, proj, tanfov, bg, H, W).forward()
        ctx.nv = nv
        rb = nv.rb
        return rb.rgb[0], rb.normal[0], rb.depth[0], rb.opacity[0], rb.confidence[0]

    @staticmethod
    def backward(ctx, d_rgb, d_normal, d_depth, d_opacity, d_conf):
        un = lambda t: None if t is None else t.unsqueeze(0)
        dm, ds, dr, do, dc = ctx.nv.backward(un(d_rgb), un(d_normal), un(d_depth), un(d_opacity), un(d_conf))
        return dm, dc, do, ds, dr, None, None


def render_views(attrs, extrinsics, intrinsics, bg, near_far, H, W):
    """utils/operations.py:829-904 with require_grad=True: sequential loop over the batch, then the
    post-processing of render_cuda_core (:714-719) in torch."""
    means, harmonics, opac, conf, scales, rots = attrs
    dev = means.device
    fovs, view, proj, _, tanfov = O.camera_blocks(extrinsics.detach().float().cpu(), intrinsics.detach().float().cpu(),
                                                  near_far)
    view, proj, tanfov, fovs = view.to(dev), proj.to(dev), tanfov.to(dev), fovs.to(dev)
    colors = harmonics[:, 0, :]
    outs = [_NaiveRasterize.apply(means, colors, opac, scales, rots, conf, (view[i], proj[i], tanfov[i], bg, H, W))
            for i in range(extrinsics.shape[0])]
    rgb, normal, depth, opacity, confidence = [torch.stack([o[k] for o in outs]) for k in range(5)]
    m = opacity.detach() > 1e-2
    normal_u = F.normalize(normal, dim=1) * m
    d2n = O._depth2normal_torch(depth, m, fovs)
    return rgb, depth, normal_u, opacity, d2n, confidence


def _one_sided_sq_diffs(x):
    """(b,c,h,w) -> (b,4,h,w): squared norms of the left/right/up/down one-sided differences, zero where the
    neighbour is outside (mapping/utils.py:44-62)"""
    dl = F.pad(x[:, :, :, :-1] - x[:, :, :, 1:], (0, 1, 0, 0))
    dr = F.pad(x[:, :, :, 1:] - x[:, :, :, :-1], (1, 0, 0, 0))
    du = F.pad(x[:, :, :-1, :] - x[:, :, 1:, :], (0, 0, 0, 1))
    dd = F.pad(x[:, :, 1:, :] - x[:, :, :-1, :], (0, 0, 1, 0))
    return (torch.stack([dl, dr, du, dd], dim=2) ** 2).sum(dim=1)


def train_loss(rgb, depth, normal, opacity, d2n, rgb_gt, depth_gt, sigma=0.3):
    """mapping/gaussian_map.py:106-124 as ATen ops (unfused), quirk Q1 (the (B,B,H,W) broadcast) included.
    Returns (total, per-frame rgb-L1 + depth-L1)."""
    m_vis = opacity.detach() > 1e-3
    m_d = depth_gt > 0.0
    l_rgb = torch.abs((rgb - rgb_gt) * m_vis)
    l_d = torch.abs((depth - depth_gt) * m_d)
    perf = l_rgb.mean(dim=[1, 2, 3]).detach() + l_d.mean(dim=[1, 2, 3]).detach()
    nd, dd = _one_sided_sq_diffs(normal), _one_sided_sq_diffs(depth.detach())
    tv = torch.mean((dd <= 1e-4).float() * torch.exp(-nd / (2 * sigma ** 2)) * nd * m_d)
    cons = 1 - torch.sum(normal * d2n, 1)
    cons = (cons * m_vis.long()).mean()
    return l_rgb.mean() + 0.8 * l_d.mean() + 0.1 * cons + 0.1 * tv, perf


class NaiveTrainer:
    """The reference's train() loop on the lineage rasterizer (fresh torch Adam, 5 groups)."""

    def __init__(self, state, frames, cfg, dev):
        self.dev, self.cfg = dev, cfg
        p = lambda k: state[k].clone().to(dev).float().requires_grad_(True)
        self.means, self.scales, self.rots = p("means"), p("scales"), p("rotations")
        self.opac, self.harm = p("opacities"), p("harmonics")
        self.view_scores = state["view_scores"].to(dev)
        self.view_means = state["view_means"].to(dev)
        self.frames = [{k: (v.to(dev) if torch.is_tensor(v) and k in ("rgb", "depth") else v) for k, v in f.items()}
                       for f in frames]
        o = cfg.optimizer
        self.opt = torch.optim.Adam([
            {"params": [self.means], "lr": o.mean_lr}, {"params": [self.scales], "lr": o.scale_lr},
            {"params": [self.rots], "lr": o.rotation_lr}, {"params": [self.opac], "lr": o.opacity_lr},
            {"params": [self.harm], "lr": o.harmonic_lr}], eps=1e-15)
        self.bg = torch.tensor(cfg.background, dtype=torch.float32, device=dev)
        self.near_far = tuple(cfg.bound)
        self.scale_factor = cfg.scale_factor
        self.instances = 0

    def attrs(self):
        vv = self.view_means.norm(dim=-1)
        vv = torch.where(torch.isnan(vv), torch.ones_like(vv), vv)
        conf = torch.clamp(torch.exp(1 - vv) * self.view_scores, min=0, max=1)
        return (self.means, self.harm, torch.sigmoid(self.opac), conf,
                torch.clamp(self.scale_factor * torch.exp(self.scales), min=0, max=0.05), F.normalize(self.rots))

    def step(self, ids):
        fr = [self.frames[int(i)] for i in ids]
        rgb_gt = torch.stack([f["rgb"] for f in fr])
        depth_gt = torch.stack([f["depth"] for f in fr])
        ext = torch.stack([f["extrinsic"] for f in fr])
        K = torch.stack([f["intrinsic"] for f in fr])
        _, H, W = fr[0]["rgb"].shape
        rgb, depth, normal, opacity, d2n, _ = render_views(self.attrs(), ext, K, self.bg, self.near_far, H, W)
        loss, perf = train_loss(rgb, depth, normal, opacity, d2n, rgb_gt, depth_gt)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss, perf
