"""Thin Python wrappers over the non-rasterizer entry points of libags_b200.so:
the fused image loss (K8) and the fused Adam (K7)."""
import ctypes as C
import torch

from . import lib as L


class LossResult:
    __slots__ = ("normal_unit", "d2n", "d_rgb", "d_normal", "d_depth", "terms", "workspace", "args", "keep")

    def total(self, w_depth=0.8, w_cons=0.1, w_tv=0.1):
        t = self.terms
        return t[0] + w_depth * t[1] + w_cons * t[2] + w_tv * t[3]

    def frame_perf(self):
        """what track_performance stores: per-frame rgb-L1 mean + depth-L1 mean"""
        return self.terms[4::2] + self.terms[5::2]


def loss_forward_backward(rgb, normal, depth, opacity, rgb_gt, depth_gt, tanfov, *, B_total=None,
                          vis_count=None, w_depth=0.8, w_cons=0.1, w_tv=0.1, out=None):
    """Post-processing + loss + gradients for a batch of rendered frames (all (B,C,H,W) CUDA fp32).
    `tanfov` is (B,2) tan(fov/2).  Returns a LossResult; `terms` = [L_rgb, L_depth, L_cons, L_tv,
    (rgb_f, depth_f) per frame].  Passing `out` (a previous result) reuses all buffers and the cached
    argument struct (the training loop calls this with fixed pointers every iteration)."""
    lib = L.load()
    if out is not None and getattr(out, "args", None) is not None:
        out.args.stream = L.current_stream(rgb.device)
        L.check(lib.ags_loss_forward_backward(C.byref(out.args)), "ags_loss_forward_backward")
        return out
    B, _, H, W = rgb.shape
    dev = rgb.device
    r = LossResult()
    o = dict(device=dev, dtype=torch.float32)
    r.normal_unit = torch.empty(B, 3, H, W, **o)
    r.d2n = torch.empty(B, 3, H, W, **o)
    r.d_rgb = torch.empty(B, 3, H, W, **o)
    r.d_normal = torch.empty(B, 3, H, W, **o)
    r.d_depth = torch.empty(B, 1, H, W, **o)
    r.terms = torch.empty(4 + 2 * B, **o)
    r.workspace = torch.empty(lib.ags_loss_scratch_bytes(B, H, W), device=dev, dtype=torch.uint8)
    a = L.LossArgs()
    a.B, a.H, a.W = B, H, W
    a.B_total = int(B_total or B)
    a.rgb, a.normal, a.depth, a.opacity = L.ptr(rgb), L.ptr(normal), L.ptr(depth), L.ptr(opacity)
    a.rgb_gt, a.depth_gt, a.tanfov = L.ptr(rgb_gt), L.ptr(depth_gt), L.ptr(tanfov)
    a.vis_count = L.ptr(vis_count)
    a.normal_unit, a.d2n = L.ptr(r.normal_unit), L.ptr(r.d2n)
    a.d_rgb, a.d_normal, a.d_depth = L.ptr(r.d_rgb), L.ptr(r.d_normal), L.ptr(r.d_depth)
    a.loss_terms = L.ptr(r.terms)
    a.w_depth, a.w_cons, a.w_tv = w_depth, w_cons, w_tv
    a.workspace, a.workspace_bytes = L.ptr(r.workspace), r.workspace.numel()
    a.stream = L.current_stream(dev)
    r.args = a
    r.keep = (rgb, normal, depth, opacity, rgb_gt, depth_gt, tanfov, vis_count)   # pointers stay valid
    L.check(lib.ags_loss_forward_backward(C.byref(a)), "ags_loss_forward_backward")
    return r


def adam_step(params, grads, exp_avgs, exp_avg_sqs, lrs, step=None, step_dev=None,
              betas=(0.9, 0.999), eps=1e-15, skip_flag_ptr=None, cache=None):
    """One fused Adam step over up to 5 groups (in place on params / exp_avg / exp_avg_sq).
    `cache` (a dict) keeps the argument struct between calls with identical tensors."""
    lib = L.load()
    if cache is not None and "a" in cache:
        a = cache["a"]
        a.step = int(step or 0)
        a.stream = L.current_stream(params[0].device)
        L.check(lib.ags_adam_step(C.byref(a)), "ags_adam_step")
        return
    a = L.AdamArgs()
    n = len(params)
    a.num_groups = n
    for k in range(n):
        a.param[k], a.grad[k] = L.ptr(params[k]), L.ptr(grads[k])
        a.exp_avg[k], a.exp_avg_sq[k] = L.ptr(exp_avgs[k]), L.ptr(exp_avg_sqs[k])
        a.numel[k] = params[k].numel()
        a.lr[k] = lrs[k]
    a.beta1, a.beta2, a.eps = betas[0], betas[1], eps
    a.step = int(step or 0)
    a.step_dev = L.ptr(step_dev)
    a.skip_flag = skip_flag_ptr
    a.stream = L.current_stream(params[0].device)
    if cache is not None:
        cache["a"] = a
    L.check(lib.ags_adam_step(C.byref(a)), "ags_adam_step")


def postprocess(normal, depth, opacity, tanfov):
    """normalize(normal)*mask and depth2normal for B rendered views (forward only); tanfov (B,2)."""
    lib = L.load()
    B, _, H, W = normal.shape
    nu = torch.empty_like(normal)
    d2n = torch.empty_like(normal)
    L.check(lib.ags_postprocess(B, H, W, L.ptr(normal), L.ptr(depth), L.ptr(opacity), L.ptr(tanfov),
                                L.ptr(nu), L.ptr(d2n), L.current_stream(normal.device)),
            "ags_postprocess")
    return nu, d2n


def smooth_depth(depth, d=15, sigma_color=0.5, sigma_space=20.0):
    """get_smooth_depth (utils/operations.py:161-169) on the device: (H,W) or (1,H,W) CUDA float32
    depth with invalid < 0 -> bilateral-filtered depth, -1 at invalid pixels."""
    lib = L.load()
    shape = depth.shape
    H, W = shape[-2], shape[-1]
    src = depth.reshape(H, W).to(torch.float32).contiguous()
    out = torch.empty_like(src)
    scratch = torch.empty(4, dtype=torch.int32, device=src.device)
    L.check(lib.ags_smooth_depth(H, W, L.ptr(src), L.ptr(out), d, sigma_color, sigma_space, L.ptr(scratch),
                                 L.current_stream(src.device)), "ags_smooth_depth")
    return out.reshape(shape)
