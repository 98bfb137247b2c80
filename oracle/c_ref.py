"""ctypes wrapper of the plain-C/OpenMP oracle (oracle/c/ags_ref.c) -- TEST INFRASTRUCTURE ONLY.

`rasterize(...)` has the keyword surface of oracle.rasterizer_ref.rasterize and is differentiable
through a torch.autograd.Function whose backward is the C hand-derived backward.  Used by
tests/test_oracle_c.py (cross-check of the two oracles) and by bench.py's cpu_baseline /
`--impl reference` arm (the reference has no CPU implementation of this path: DESIGN.md section 1).
PARITY UNPINNED for the native half, see oracle/rasterizer_ref.py.
"""
import ctypes as C
import os
import subprocess
import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def _lib(dtype):
    key = "f64" if dtype == torch.float64 else "f32"
    if key not in _libs:
        path = os.path.join(_HERE, f"libags_ref_{key}.so")
        src = os.path.join(_HERE, "c", "ags_ref.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        lib = C.CDLL(path)
        lib.agsref_forward.restype = C.c_void_p
        lib.agsref_forward.argtypes = [C.c_void_p]
        lib.agsref_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.agsref_free.argtypes = [C.c_void_p]
        lib.agsref_num_instances.argtypes = [C.c_void_p]
        lib.agsref_set_threshold_shift.argtypes = [C.c_double]
        lib.agsref_set_threshold_shift.restype = None
        assert lib.agsref_sizeof_real() == (8 if key == "f64" else 4)
        _libs[key] = lib
    return _libs[key]


def _structs(real):
    P = C.c_void_p

    class Fwd(C.Structure):
        _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("require_importance", C.c_int32),
                    ("front_only", C.c_int32), ("tanfovx", real), ("tanfovy", real), ("scale_modifier", real),
                    ("weight_thres", real)] + [(n, P) for n in (
                        "means", "scales", "rots", "opac", "colors", "conf", "view", "proj", "bg", "mask",
                        "rgb", "normal", "depth", "opacity", "confidence", "importance", "count", "radii")]

    class Bwd(C.Structure):
        _fields_ = [(n, P) for n in ("d_rgb", "d_normal", "d_depth", "d_opacity", "d_conf", "d_means",
                                     "d_means2d", "d_opac", "d_colors", "d_scales", "d_rots")]
    return Fwd, Bwd


class _RasterC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, colors, scales, rotations, confidences, cfg):
        dt = means3D.dtype
        lib = _lib(dt)
        real = C.c_double if dt == torch.float64 else C.c_float
        Fwd, Bwd = _structs(real)
        N, H, W = means3D.shape[0], cfg["H"], cfg["W"]
        c = lambda t: t.detach().to(dt).contiguous()
        keep = dict(means=c(means3D), scales=c(scales), rots=c(rotations), opac=c(opacities).reshape(-1),
                    colors=c(colors), conf=c(confidences), view=c(cfg["viewmatrix"]).reshape(16),
                    proj=c(cfg["projmatrix"]).reshape(16), bg=c(cfg["bg"])[:3].contiguous())
        mask = cfg["render_mask"]
        keep["mask"] = c(mask).reshape(-1) if (mask is not None and mask.numel() > 0) else None
        out = dict(rgb=torch.empty(3, H, W, dtype=dt), normal=torch.empty(3, H, W, dtype=dt),
                   depth=torch.empty(1, H, W, dtype=dt), opacity=torch.empty(1, H, W, dtype=dt),
                   confidence=torch.empty(1, H, W, dtype=dt), importance=torch.zeros(N, dtype=dt),
                   count=torch.zeros(N, dtype=torch.int32), radii=torch.zeros(N, dtype=torch.int32))
        a = Fwd()
        a.N, a.H, a.W = N, H, W
        a.require_importance, a.front_only = int(cfg["require_importance"]), int(cfg["front_only"])
        a.tanfovx, a.tanfovy = cfg["tanfovx"], cfg["tanfovy"]
        a.scale_modifier, a.weight_thres = cfg["scale_modifier"], cfg["weight_thres"]
        for k, t in list(keep.items()) + list(out.items()):
            setattr(a, k, None if t is None else t.data_ptr())
        lib.agsref_set_threshold_shift(float(cfg.get("threshold_shift", 0.0)))
        try:
            state = lib.agsref_forward(C.byref(a))
        finally:
            lib.agsref_set_threshold_shift(0.0)
        ctx.n_inst = lib.agsref_num_instances(state)
        if not any(ctx.needs_input_grad):            # no backward can follow: release the oracle state now
            lib.agsref_free(state)
            state = None
        ctx.pack = (lib, a, state, keep, out, Bwd, dt, N)
        ctx.shift = float(cfg.get("threshold_shift", 0.0))
        ctx.mark_non_differentiable(out["importance"], out["count"], out["radii"])
        return (out["rgb"], out["normal"], out["depth"], out["opacity"], out["confidence"],
                out["importance"].float(), out["count"], out["radii"])

    @staticmethod
    def backward(ctx, d_rgb, d_normal, d_depth, d_opacity, d_conf, *_):
        lib, a, state, keep, out, Bwd, dt, N = ctx.pack
        g = Bwd()
        ups = [None if t is None else t.detach().to(dt).contiguous() for t in (d_rgb, d_normal, d_depth, d_opacity, d_conf)]
        for n, t in zip(("d_rgb", "d_normal", "d_depth", "d_opacity", "d_conf"), ups):
            setattr(g, n, None if t is None else t.data_ptr())
        res = dict(d_means=torch.empty(N, 3, dtype=dt), d_means2d=torch.empty(N, 3, dtype=dt),
                   d_opac=torch.empty(N, dtype=dt), d_colors=torch.empty(N, 3, dtype=dt),
                   d_scales=torch.empty(N, 3, dtype=dt), d_rots=torch.empty(N, 4, dtype=dt))
        for k, t in res.items():
            setattr(g, k, t.data_ptr())
        lib.agsref_set_threshold_shift(float(ctx.shift))
        try:
            lib.agsref_backward(C.byref(a), state, C.byref(g))
        finally:
            lib.agsref_set_threshold_shift(0.0)
        return (res["d_means"], res["d_means2d"], res["d_opac"].reshape(-1, 1), res["d_colors"], res["d_scales"],
                res["d_rots"], None, None)

    @staticmethod
    def release(ctx):
        lib, a, state = ctx.pack[:3]
        if state is not None:
            lib.agsref_free(state)


def rasterize(means3D, means2D, opacities, confidences, colors, scales, rotations, *, image_height,
              image_width, tanfovx, tanfovy, bg, viewmatrix, projmatrix, scale_modifier=1.0,
              render_mask=None, weight_thres=0.03, require_importance=False, front_only=False,
              threshold_shift=0.0):
    """`threshold_shift` (tests only): relative shift of every hard threshold of the specification
    (alpha cut-offs, T stop, radius rounding, tile-rect truncation, culls, clamp edges)."""
    cfg = dict(threshold_shift=float(threshold_shift), H=int(image_height), W=int(image_width), tanfovx=float(tanfovx), tanfovy=float(tanfovy), bg=bg,
               viewmatrix=viewmatrix, projmatrix=projmatrix, scale_modifier=float(scale_modifier),
               render_mask=render_mask, weight_thres=float(weight_thres),
               require_importance=require_importance, front_only=front_only)
    if means2D is None:
        means2D = torch.zeros_like(means3D)
    return _RasterC.apply(means3D, means2D, opacities.reshape(-1, 1), colors, scales, rotations, confidences, cfg)


def forward_backward(attrs, view, proj, tanfov, hw, ups=None, *, dtype=torch.float64, bg=None, render_mask=None,
                     require_importance=False, front_only=False, weight_thres=0.03, threshold_shift=0.0):
    """One view through the C oracle without autograd bookkeeping: `attrs` = (means, colors (N,3),
    opacities (N,), confidences, scales, rotations) ACTIVATED tensors, `ups` = upstream gradients of
    (rgb, normal, depth, opacity, confidence) or None.  Returns (outputs (8), grads (6) or None) with
    grads ordered (means3D, means2D, opacity, colors, scales, rotations); the oracle state is freed."""
    means, colors, opac, conf, scales, rots = [t.detach().to(dtype) for t in attrs]
    leaf = lambda t: t.clone().requires_grad_(ups is not None)
    m, o, c, s, r = leaf(means), leaf(opac.reshape(-1, 1)), leaf(colors), leaf(scales), leaf(rots)
    m2 = torch.zeros_like(m, requires_grad=ups is not None)
    H, W = hw
    out = rasterize(m, m2, o, conf, c, s, r, image_height=H, image_width=W, tanfovx=float(tanfov[0]),
                    tanfovy=float(tanfov[1]), bg=torch.zeros(4) if bg is None else bg, viewmatrix=view.to(dtype),
                    projmatrix=proj.to(dtype), render_mask=render_mask, weight_thres=weight_thres,
                    require_importance=require_importance, front_only=front_only, threshold_shift=threshold_shift)
    grads = None
    fn = out[0].grad_fn
    if ups is not None:
        loss = sum((u.to(dtype) * x).sum() for u, x in zip(ups, out[:5]))
        loss.backward()
        grads = [m.grad, m2.grad, o.grad.reshape(-1), c.grad, s.grad, r.grad]
    outs = [t.detach() for t in out]
    if fn is not None and hasattr(fn, "pack"):
        _RasterC.release(fn)
        fn.pack = None
    return outs, grads
