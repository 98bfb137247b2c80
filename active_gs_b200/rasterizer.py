"""Drop-in for the reference's native rasterizer module.

`GaussianRasterizationSettings` / `GaussianRasterizer` reproduce the surface that
/root/reference/utils/operations.py:22-25,682-713 imports and calls (15 settings fields, 9 call
kwargs, 8 returned tensors), so `render_cuda_core` runs unchanged on top of libags_b200.so.
`RenderBatch` is the B-view form used by the fused training loop (active_gs_b200.gaussian_map).
"""
from typing import NamedTuple
import ctypes as C
import torch

from . import lib as L


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    render_mask: torch.Tensor
    weight_thres: float
    debug: bool
    config: torch.Tensor


_cap_hint = {}          # device -> instances-per-Gaussian estimate that was enough last time
INT32_MAX = 2 ** 31 - 1


def views_per_chunk(N, H, W, budget_bytes=2 << 30, per_gaussian=None, device_index=0, cap=128):
    """How many views one RenderBatch may take so that its workspace stays within `budget_bytes`
    (planner candidates and the post-processing re-render come in chunks of this many views).  The
    per-pair part of the workspace is 140 B, an instance costs 20 B (csrc/ags_common.cuh)."""
    per = per_gaussian if per_gaussian is not None else _cap_hint.get(device_index, 4.0)
    per_view = 140 * max(N, 1) + 20 * per * max(N, 1) + 8 * H * W + 4096
    return int(max(1, min(cap, budget_bytes // per_view, (INT32_MAX - 1) // max(N, 1))))


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class RenderBatch:
    """One forward (and optionally backward) of B views through the C ABI.  Owns the workspace
    tensor, which carries the state the backward needs."""

    def __init__(self, means3D, scales, rotations, opacities, colors, confidences, viewmatrix,
                 projmatrix, tanfov, bg, H, W, *, render_mask=None, scale_modifier=1.0,
                 weight_thres=0.03, require_importance=False, front_only=False,
                 param_mode=L.PARAMS_ACTIVATED, scale_factor=0.01, scale_max=0.05, inst_cap=None,
                 with_importance=True, pool=None, images=None):
        lib = L.load()
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("active_gs_b200 rasterizer needs CUDA tensors (there is no CPU path)")
        self.dev = dev
        self.N = int(means3D.shape[0])
        self.B = int(viewmatrix.shape[0])
        self.H, self.W = int(H), int(W)
        N, B = self.N, self.B
        self.inputs = [_f32c(means3D), _f32c(scales), _f32c(rotations), _f32c(opacities).reshape(-1),
                       _f32c(colors), _f32c(confidences) if confidences is not None else None]
        self.view = [_f32c(viewmatrix).reshape(B, 16), _f32c(projmatrix).reshape(B, 16),
                     _f32c(tanfov).reshape(B, 2), _f32c(bg)]
        if self.view[3].numel() < 3:
            raise ValueError("bg needs at least 3 channels")
        self.mask = None
        if render_mask is not None and render_mask.numel() > 0:
            self.mask = _f32c(render_mask).reshape(B, self.H, self.W)
        o = dict(device=dev, dtype=torch.float32)
        if images is not None:           # preallocated (B,C,H,W) outputs of a persistent training engine
            self.rgb, self.normal, self.depth, self.opacity, self.confidence = images
        else:
            self.rgb = torch.empty(B, 3, H, W, **o)
            self.normal = torch.empty(B, 3, H, W, **o)
            self.depth = torch.empty(B, 1, H, W, **o)
            self.opacity = torch.empty(B, 1, H, W, **o)
            self.confidence = torch.empty(B, 1, H, W, **o)
        # optional outputs (all-zero unless require_importance): the training engine skips them
        with_importance = with_importance or require_importance
        self.importance = torch.empty(B, N, **o) if with_importance else None
        self.count = torch.empty(B, N, device=dev, dtype=torch.int32) if with_importance else None
        self.radii = torch.empty(B, N, device=dev, dtype=torch.int32)
        self.stats = torch.empty(L.AGS_NUM_STATS, device=dev, dtype=torch.int32)
        self.cfg = dict(param_mode=param_mode, require_importance=int(require_importance),
                        front_only=int(front_only), scale_modifier=float(scale_modifier),
                        weight_thres=float(weight_thres), scale_factor=float(scale_factor),
                        scale_max=float(scale_max))
        if inst_cap is None:
            per = _cap_hint.get(dev.index, 4.0)
            inst_cap = int(per * N * B) + 4096
        self.inst_cap = min(int(inst_cap), INT32_MAX)          # AgsRenderArgs.inst_cap is an int32
        self.workspace = None
        self.pool = pool                 # optional callable(nbytes) -> uint8 tensor with >= nbytes (reused across calls)
        self._alloc(lib)

    def _alloc(self, lib):
        nbytes = lib.ags_scratch_bytes(self.N, self.B, self.H, self.W, self.inst_cap)
        if self.pool is not None:
            self.workspace = self.pool(nbytes + 256)
        else:
            self.workspace = torch.empty(nbytes + 256, device=self.dev, dtype=torch.uint8)
        self.ws_ptr = (self.workspace.data_ptr() + 255) & ~255
        self.ws_bytes = nbytes

    def _args(self):
        a = L.RenderArgs()
        a.N, a.B, a.H, a.W = self.N, self.B, self.H, self.W
        a.param_mode = self.cfg["param_mode"]
        a.require_importance = self.cfg["require_importance"]
        a.front_only = self.cfg["front_only"]
        a.inst_cap = self.inst_cap
        a.scale_modifier = self.cfg["scale_modifier"]
        a.weight_thres = self.cfg["weight_thres"]
        a.scale_factor = self.cfg["scale_factor"]
        a.scale_max = self.cfg["scale_max"]
        (a.means3D, a.scales, a.rotations, a.opacities, a.colors, a.confidences) = [
            L.ptr(t) for t in self.inputs]
        a.viewmatrix, a.projmatrix, a.tanfov, a.bg = [L.ptr(t) for t in self.view]
        a.render_mask = L.ptr(self.mask)
        a.out_rgb, a.out_normal, a.out_depth = L.ptr(self.rgb), L.ptr(self.normal), L.ptr(self.depth)
        a.out_opacity, a.out_confidence = L.ptr(self.opacity), L.ptr(self.confidence)
        a.importance, a.count, a.radii = L.ptr(self.importance), L.ptr(self.count), L.ptr(self.radii)
        a.stats = L.ptr(self.stats)
        a.workspace, a.workspace_bytes = self.ws_ptr, self.ws_bytes
        a.stream = L.current_stream(self.dev)
        return a

    def forward(self, check_overflow=True):
        """Enqueue the forward.  With check_overflow the instance statistics are read back (one host
        sync, like the reference's own per-view .item() calls) and the forward is re-run with a
        larger workspace if the batch needed more instances than `inst_cap`."""
        lib = L.load()
        L.check(lib.ags_render_forward(C.byref(self._args())), "ags_render_forward")
        self.last_instances = 0
        if check_overflow:
            st = self.stats.tolist()
            need = st[L.STAT_INSTANCES]
            self.last_instances = need
            if st[L.STAT_OVERFLOW]:
                self.inst_cap = min(int(need * 1.25) + 4096, INT32_MAX)
                self._alloc(lib)
                L.check(lib.ags_render_forward(C.byref(self._args())), "ags_render_forward")
            if self.N * self.B > 0:
                _cap_hint[self.dev.index] = max(_cap_hint.get(self.dev.index, 4.0) * 0.9,
                                                1.5 * need / (self.N * self.B), 1.0)
        return self

    def backward(self, d_rgb=None, d_normal=None, d_depth=None, d_opacity=None, d_confidence=None,
                 want_means2D=False):
        lib = L.load()
        N, dev = self.N, self.dev
        o = dict(device=dev, dtype=torch.float32)
        g = L.RenderGradArgs()
        ups = [None if t is None else _f32c(t) for t in (d_rgb, d_normal, d_depth, d_opacity, d_confidence)]
        g.d_rgb, g.d_normal, g.d_depth, g.d_opacity, g.d_confidence = [L.ptr(t) for t in ups]
        dm, ds, dr = torch.empty(N, 3, **o), torch.empty(N, 3, **o), torch.empty(N, 4, **o)
        do, dc = torch.empty(N, **o), torch.empty(N, 3, **o)
        dm2 = torch.empty(self.B, N, 3, **o) if want_means2D else None
        g.d_means3D, g.d_scales, g.d_rotations = L.ptr(dm), L.ptr(ds), L.ptr(dr)
        g.d_opacities, g.d_colors, g.d_means2D = L.ptr(do), L.ptr(dc), L.ptr(dm2)
        g.accumulate = 0
        g.clear_records = 1
        L.check(lib.ags_render_backward(C.byref(self._args()), C.byref(g)), "ags_render_backward")
        return dm, ds, dr, do, dc, dm2


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, confidences, colors, scales, rotations, settings):
        s = settings
        cfg = s.config.detach().float().cpu().tolist() if s.config is not None else [1, 1, 1, 0, 0]
        dev = means3D.device
        tanfov = torch.tensor([[s.tanfovx, s.tanfovy]], dtype=torch.float32, device=dev)
        rb = RenderBatch(
            means3D, scales, rotations, opacities, colors, confidences,
            s.viewmatrix.reshape(1, 4, 4), s.projmatrix.reshape(1, 4, 4), tanfov, s.bg,
            s.image_height, s.image_width, render_mask=s.render_mask,
            scale_modifier=s.scale_modifier, weight_thres=s.weight_thres,
            require_importance=cfg[3] > 0, front_only=cfg[4] > 0)
        rb.forward(check_overflow=True)
        ctx.rb = rb
        ctx.opac_shape = opacities.shape
        ctx.mark_non_differentiable(rb.importance, rb.count, rb.radii)
        return (rb.rgb[0], rb.normal[0], rb.depth[0], rb.opacity[0], rb.confidence[0],
                rb.importance[0], rb.count[0], rb.radii[0])

    @staticmethod
    def backward(ctx, d_rgb, d_normal, d_depth, d_opacity, d_conf, *_):
        rb = ctx.rb
        un = lambda t: None if t is None else t.unsqueeze(0)
        dm, ds, dr, do, dc, dm2 = rb.backward(un(d_rgb), un(d_normal), un(d_depth), un(d_opacity),
                                              un(d_conf), want_means2D=True)
        # ctx.rb stays: the backward consumed-and-cleared the gradient records (clear_records=1), so a
        # second backward over the same graph (retain_graph=True) starts from zero again
        return dm, dm2[0], do.reshape(ctx.opac_shape), None, dc, ds, dr, None


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, confidences=None, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        if shs is not None or colors_precomp is None:
            raise ValueError("active_gs_b200: only colors_precomp is supported "
                             "(the reference passes sh_degree=0, shs=None)")
        if cov3D_precomp is not None or scales is None or rotations is None:
            raise ValueError("active_gs_b200: scales/rotations are required (cov3D_precomp unsupported)")
        if confidences is None:
            confidences = torch.zeros(means3D.shape[0], device=means3D.device)
        return _Rasterize.apply(means3D, means2D, opacities, confidences, colors_precomp, scales,
                                rotations, self.raster_settings)
