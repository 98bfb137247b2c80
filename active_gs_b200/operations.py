"""Host-side mirror of the hot-path part of the reference's utils/operations.py, on top of
libags_b200.so.  Same names, argument meaning and return order as the reference so that
mapping/mapper.py, the planners, eval and the GUI can call it unchanged:

    GaussianRenderer(extrinsics, intrinsics, gaussians_attr, background_color, near_far,
                     resolution, device, render_masks=None)
        .render_view(i, require_grad, require_importance, front_only)
        .render_view_all(require_grad, require_importance, front_only)
    -> (rgb, depth, normal, opacity, d2n, confidence, importance, count, in_frustum_mask)

B200-first differences (results identical): all B views are rendered by ONE library call
(reference: Python loop utils/operations.py:853-892), the per-view .item()/H2D syncs
(:685-686,697) are gone (tan(fov/2) travels as a (B,2) device tensor), and the post-processing
(:714-718) is one fused kernel instead of ~30 ATen launches per view.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .rasterizer import RenderBatch


def inverse_sigmoid(x):
    """utils/operations.py:101-102."""
    return torch.log(x / (1 - x))


def fov2focal(fov, pixels):
    """utils/operations.py:157-158."""
    return pixels / (2 * math.tan(fov / 2))


def get_smooth_depth(depth, tolerance=0.5):
    """utils/operations.py:161-169: bilateral filter (d=15, sigmaColor=tolerance, sigmaSpace=20)
    of the valid depth on the CPU, invalid pixels set back to -1."""
    import cv2
    invalid = depth < 0.0
    d = np.where(invalid, 0.0, depth).astype(np.float32)
    out = cv2.bilateralFilter(d, 15, tolerance, 20)
    out[invalid] = -1.0
    return out


def get_smooth_depth_device(depth, tolerance=0.5):
    """get_smooth_depth for a CUDA depth tensor ((H,W) or (1,H,W)): the same OpenCV bilateral filter
    (d=15, sigmaColor=tolerance, sigmaSpace=20), restated as a CUDA kernel -- no D2H/H2D round trip
    (utils/operations.py:161-169 runs cv2 on the CPU: ~30 ms per keyframe at 640x480)."""
    return ops.smooth_depth(depth, d=15, sigma_color=tolerance, sigma_space=20.0)


def quaternion_to_matrix(q):
    """utils/operations.py:261-278, (r,x,y,z)."""
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)],
                       -1).reshape(len(q), 3, 3)


def normal2rotation(n):
    """utils/operations.py:481-541 -> (quaternion, rotation matrix)."""
    z = n / n.norm(dim=1, keepdim=True)
    ref = torch.zeros_like(z)
    ref[:, 0] = 1.0
    ref[z[:, 0].abs() > 0.99] = torch.tensor([0.0, 1.0, 0.0], device=z.device)
    x = ref - (ref * z).sum(1, keepdim=True) * z
    x = x / x.norm(dim=1, keepdim=True)
    y = torch.linalg.cross(z, x)
    y = y / y.norm(dim=1, keepdim=True)
    R = torch.stack([x, y, z], -1)
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2] + 1e-6
    r = torch.sqrt(1 + tr) / 2
    q = torch.stack([r, (R[:, 2, 1] - R[:, 1, 2]) / (4 * r), (R[:, 0, 2] - R[:, 2, 0]) / (4 * r),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * r)], -1)
    return F.normalize(q, dim=-1), R


def get_fov(intrinsics):
    """utils/operations.py:628-642."""
    Kinv = torch.linalg.inv(intrinsics)

    def ray(u, v):
        d = Kinv @ torch.tensor([u, v, 1.0], dtype=intrinsics.dtype, device=intrinsics.device)
        return d / d.norm(dim=-1, keepdim=True)

    fx = (ray(0.0, 0.5) * ray(1.0, 0.5)).sum(-1).acos()
    fy = (ray(0.5, 0.0) * ray(0.5, 1.0)).sum(-1).acos()
    return torch.stack([fx, fy], -1)


def get_projection_matrix(near, far, fov_x, fov_y):
    """utils/operations.py:572-600."""
    tx, ty = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    P = torch.zeros(near.shape[0], 4, 4, dtype=torch.float32, device=near.device)
    right, top = tx * near, ty * near
    P[:, 0, 0] = 2 * near / (right + right)
    P[:, 1, 1] = 2 * near / (top + top)
    P[:, 3, 2] = 1
    P[:, 2, 2] = far / (far - near)
    P[:, 2, 3] = -(far * near) / (far - near)
    return P


def camera_blocks(extrinsics, intrinsics, near_far):
    """utils/operations.py:748-762 on whatever device the inputs live on (CPU is fine: 4x4 math).
    Returns fovs (B,2), viewmatrix (B,4,4), projmatrix (B,4,4), campos (B,3), tanfov (B,2)."""
    B = extrinsics.shape[0]
    dev = extrinsics.device
    near = torch.full((B,), float(near_far[0]), device=dev)
    far = torch.full((B,), float(near_far[1]), device=dev)
    fovs = get_fov(intrinsics)
    Pt = get_projection_matrix(near, far, fovs[:, 0], fovs[:, 1]).transpose(1, 2)
    view = torch.linalg.inv(extrinsics).transpose(1, 2)
    return fovs, view, view @ Pt, extrinsics[:, :3, 3], (0.5 * fovs).tan()


def unproject_grid(H, W, intrinsics, device):
    """utils/operations.py:372-392,464-478: camera-space rays (z=1) through the pixel centres,
    (H*W, 3) for one normalised K."""
    ys = (torch.arange(H, dtype=torch.float32, device=device) + 0.5) / H
    xs = (torch.arange(W, dtype=torch.float32, device=device) + 0.5) / W
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([xx, yy, torch.ones_like(xx)], -1).reshape(-1, 3)
    return pix @ torch.linalg.inv(intrinsics).t()


def get_world_rays(H, W, extrinsic, intrinsic, device):
    """utils/operations.py:544-569: (origins, directions) of all pixels, un-normalised directions."""
    d = unproject_grid(H, W, intrinsic.to(device), device) @ extrinsic[:3, :3].to(device).t()
    return extrinsic[:3, 3].to(device).expand_as(d), d


def depth2normal(depth, mask, fov):
    """utils/operations.py:172-219 for ONE (1,H,W) depth map, via the fused post-process kernel
    (quirk Q2 kept).  `mask` must be 0/1; it is applied as an opacity plane."""
    dev = depth.device
    H, W = depth.shape[1:]
    opac = mask.to(torch.float32).reshape(1, 1, H, W).contiguous()
    fv = torch.tensor([[math.tan(0.5 * float(fov[0])), math.tan(0.5 * float(fov[1]))]],
                      dtype=torch.float32, device=dev)
    _, d2n = ops.postprocess(torch.zeros(1, 3, H, W, device=dev), depth.reshape(1, 1, H, W).contiguous(),
                             opac, fv)
    return d2n[0]


def voxel_downsample(points, voxel_size=0.02):
    """utils/operations.py:603-625: indices of one randomly chosen point per occupied voxel."""
    vox = torch.floor(points / voxel_size).long()
    uniq, inv = torch.unique(vox, return_inverse=True, dim=0)
    perm = torch.randperm(inv.numel(), device=points.device)
    sel = torch.zeros(uniq.shape[0], dtype=torch.long, device=points.device)
    sel[inv[perm]] = perm
    return sel.unique()


class _RenderAll(torch.autograd.Function):
    """Autograd bridge for GaussianRenderer(require_grad=True): B views, one library call each way."""

    @staticmethod
    def forward(ctx, means, colors, opac, scales, rots, conf, r):
        rb = RenderBatch(means, scales, rots, opac, colors, conf, r._view, r._proj, r._tanfov,
                         r.background_color, r.h, r.w, render_mask=r._mask_tensor(),
                         require_importance=r._req_imp, front_only=r._front)
        rb.forward(check_overflow=True)
        ctx.rb = rb
        ctx.mark_non_differentiable(rb.importance, rb.count, rb.radii)
        return rb.rgb, rb.normal, rb.depth, rb.opacity, rb.confidence, rb.importance, rb.count, rb.radii

    @staticmethod
    def backward(ctx, d_rgb, d_normal, d_depth, d_opacity, d_conf, *_):
        dm, ds, dr, do, dc, _ = ctx.rb.backward(d_rgb, d_normal, d_depth, d_opacity, d_conf)
        return dm, dc, do, ds, dr, None, None


class GaussianRenderer:
    """utils/operations.py:723-904."""

    def __init__(self, extrinsics, intrinsics, gaussians_attr, background_color, near_far,
                 resolution, device, render_masks=None):
        self.device = torch.device(device)
        self.update_attr(gaussians_attr)
        self.background_color = background_color
        self.h, self.w = resolution
        self.batch_size = extrinsics.shape[0]
        # 4x4 camera algebra on the host (the reference does it with ~40 tiny GPU kernels)
        fovs, view, proj, campos, tanfov = camera_blocks(extrinsics.detach().float().cpu(),
                                                         intrinsics.detach().float().cpu(), near_far)
        self.cam_pos = campos.to(self.device)
        self.fovs = fovs.to(self.device)
        self.view_matrices = view.to(self.device)
        self.projection_matrices = proj.to(self.device)
        self._view, self._proj = self.view_matrices.contiguous(), self.projection_matrices.contiguous()
        self._tanfov = tanfov.to(self.device).contiguous()
        rays = unproject_grid(self.h, self.w, intrinsics[0].detach().float().cpu(), "cpu")
        self.raydir_map = F.normalize(rays, dim=-1).reshape(self.h, self.w, 3).permute(2, 0, 1).to(self.device)
        self.render_masks = render_masks
        self._req_imp = self._front = False

    def update_attr(self, gaussians_attr):
        (self.gaussian_means, self.gaussian_harmonics, self.gaussian_opacities,
         self.gaussian_confidences, self.gaussian_scales, self.gaussian_rotations) = gaussians_attr

    def _mask_tensor(self, i=None):
        if self.render_masks is None:
            return None
        m = self.render_masks if torch.is_tensor(self.render_masks) else torch.stack(list(self.render_masks))
        if m.numel() == 0:
            return None
        m = m.reshape(self.batch_size, self.h, self.w)
        return m if i is None else m[i:i + 1]

    def _render(self, sel, require_grad, require_importance, front_only):
        self._req_imp, self._front = require_importance, front_only
        view, proj, tanfov, fovs = self._view, self._proj, self._tanfov, self.fovs
        mask = self._mask_tensor()
        if sel is not None:
            view, proj, tanfov, fovs = view[sel:sel + 1], proj[sel:sel + 1], tanfov[sel:sel + 1], fovs[sel:sel + 1]
            mask = None if mask is None else mask[sel:sel + 1]
        colors = self.gaussian_harmonics[:, 0, :]
        with torch.set_grad_enabled(require_grad):
            if require_grad and torch.is_grad_enabled():
                sub = self if sel is None else _View(self, view, proj, tanfov, mask)
                rgb, normal, depth, opacity, conf, imp, cnt, radii = _RenderAll.apply(
                    self.gaussian_means, colors, self.gaussian_opacities, self.gaussian_scales,
                    self.gaussian_rotations, self.gaussian_confidences, sub)
                m = (opacity.detach() > 1e-2)
                normal_u = F.normalize(normal, dim=1) * m
                d2n = _depth2normal_torch(depth, m, fovs)
            else:
                rb = RenderBatch(self.gaussian_means, self.gaussian_scales, self.gaussian_rotations,
                                 self.gaussian_opacities, colors, self.gaussian_confidences, view, proj,
                                 tanfov, self.background_color, self.h, self.w, render_mask=mask,
                                 require_importance=require_importance, front_only=front_only)
                rb.forward(check_overflow=True)
                rgb, normal, depth, opacity, conf = rb.rgb, rb.normal, rb.depth, rb.opacity, rb.confidence
                imp, cnt, radii = rb.importance, rb.count, rb.radii
                normal_u, d2n = ops.postprocess(normal, depth, opacity, tanfov.contiguous())
        return rgb, depth, normal_u, opacity, d2n, conf, imp, cnt, radii

    def render_view(self, i=0, require_grad=False, require_importance=False, front_only=False):
        o = self._render(i, require_grad, require_importance, front_only)
        return tuple(t[0] for t in o[:8]) + (o[8][0] > 0,)

    def render_view_all(self, require_grad=False, require_importance=False, front_only=False):
        o = self._render(None, require_grad, require_importance, front_only)
        return o[:8] + (o[8].sum(0) > 0,)


class _View:
    """single-view slice of a GaussianRenderer for the autograd bridge"""

    def __init__(self, r, view, proj, tanfov, mask):
        self._view, self._proj, self._tanfov, self._mask = view, proj, tanfov, mask
        self.background_color, self.h, self.w = r.background_color, r.h, r.w
        self._req_imp, self._front = r._req_imp, r._front

    def _mask_tensor(self):
        return self._mask


def _depth2normal_torch(depth, mask, fovs):
    """Differentiable torch form of utils/operations.py:172-219 for (B,1,H,W) (only used by the
    autograd-compatibility path; the training loop uses the fused kernel)."""
    B, _, H, W = depth.shape
    dev = depth.device
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                            torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    k00 = H / (2 * torch.tan(fovs[:, 0] / 2))
    k11 = W / (2 * torch.tan(fovs[:, 1] / 2))
    d = depth[:, 0]
    pos = torch.stack([(ii - 0.5 * W) * d / k00[:, None, None], (jj - 0.5 * H) * d / k11[:, None, None], d], 1)
    m = mask.to(torch.float32)
    pp = F.pad(pos, (1, 1, 1, 1), mode="replicate")
    mp = F.pad(m, (1, 1, 1, 1), mode="replicate")
    c = pp[:, :, 1:-1, 1:-1] * mp[:, :, 1:-1, 1:-1]

    def nb(dy, dx):
        return (pp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W] - c) * mp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]

    u, l, b, r = nb(-1, 0), nb(0, -1), nb(1, 0), nb(0, 1)
    cr = lambda a, b_: torch.linalg.cross(a, b_, dim=1)
    n = F.normalize(cr(u, l) + cr(r, u) + cr(b, r) + cr(l, b), dim=1)
    return n * m
