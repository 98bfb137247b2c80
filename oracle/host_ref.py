"""CPU oracle of the HOST half of the hot path (TEST INFRASTRUCTURE ONLY).

Plain-PyTorch restatement of what ActiveGS does around the native rasterizer, each function citing
the reference lines it follows (paths relative to /root/reference).  This half IS pinned: it is
checked against fixtures produced by importing the reference's own Python
(tests/golden/make_golden.py -> tests/golden/host_golden.pt).  The native half is in
oracle/rasterizer_ref.py (parity unpinned, see its header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import rasterizer_ref as rr


# ---------------------------------------------------------------- camera set-up
def get_fov(intrinsics):
    """utils/operations.py:628-642 -- angle between the back-projected image-edge midpoints of a
    normalised K (B,3,3) -> (B,2) [fov_x, fov_y] radians."""
    Kinv = torch.linalg.inv(intrinsics)

    def ray(u, v):
        d = Kinv @ torch.tensor([u, v, 1.0], dtype=intrinsics.dtype, device=intrinsics.device)
        return d / d.norm(dim=-1, keepdim=True)

    fx = (ray(0.0, 0.5) * ray(1.0, 0.5)).sum(-1).acos()
    fy = (ray(0.5, 0.0) * ray(0.5, 1.0)).sum(-1).acos()
    return torch.stack([fx, fy], -1)


def projection_matrix(near, far, fov_x, fov_y):
    """utils/operations.py:572-600 -- z in (0,1), w = z_view. (B,) inputs -> (B,4,4)."""
    tx, ty = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    top, right = ty * near, tx * near
    P = torch.zeros(near.shape[0], 4, 4, dtype=torch.float32, device=near.device)
    P[:, 0, 0] = 2 * near / (2 * right)
    P[:, 1, 1] = 2 * near / (2 * top)
    P[:, 0, 2] = 0.0
    P[:, 1, 2] = 0.0
    P[:, 3, 2] = 1
    P[:, 2, 2] = far / (far - near)
    P[:, 2, 3] = -(far * near) / (far - near)
    return P


def camera_setup(extrinsics, intrinsics, near_far):
    """utils/operations.py:748-762 -- (B,4,4) c2w + (B,3,3) normalised K ->
    fovs (B,2), viewmatrix (B,4,4) = (w2c)^T, projmatrix (B,4,4) = viewmatrix @ P^T, campos."""
    B = extrinsics.shape[0]
    dev = extrinsics.device
    near = torch.full((B,), float(near_far[0]), device=dev)
    far = torch.full((B,), float(near_far[1]), device=dev)
    fovs = get_fov(intrinsics)
    Pt = projection_matrix(near, far, fovs[:, 0], fovs[:, 1]).transpose(1, 2)
    view = torch.linalg.inv(extrinsics).transpose(1, 2)
    return fovs, view, view @ Pt, extrinsics[:, :3, 3]


def raydir_map(intrinsics0, H, W):
    """utils/operations.py:764-772 -- unit ray directions of view 0, (3,H,W)."""
    ys = (torch.arange(H, dtype=torch.float32) + 0.5) / H
    xs = (torch.arange(W, dtype=torch.float32) + 0.5) / W
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([xx, yy, torch.ones_like(xx)], -1).to(intrinsics0.device)
    d = pix @ torch.linalg.inv(intrinsics0).t()
    return F.normalize(d, dim=-1).permute(2, 0, 1)


# ---------------------------------------------------------------- depth -> normal (quirk Q2)
def fov2focal(fov, pixels):
    """utils/operations.py:157-158."""
    return pixels / (2 * math.tan(fov / 2))


def depth2normal(depth, mask, fov):
    """utils/operations.py:172-219.  depth (1,H,W), mask (1,H,W) bool, fov = (fov_x, fov_y).
    Keeps quirk Q2: fov[0] is paired with the image HEIGHT, fov[1] with the WIDTH, and the principal
    point is (W/2, H/2) with pixel coordinates at integers."""
    H, W = depth.shape[1:]
    d = depth[0]
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=d.device),
                            torch.arange(W, dtype=torch.float32, device=d.device), indexing="ij")
    k00 = fov2focal(float(fov[0]), H)
    k11 = fov2focal(float(fov[1]), W)
    X = (ii - 0.5 * W) * d / k00
    Y = (jj - 0.5 * H) * d / k11
    pos = torch.stack([X, Y, d], -1)                                   # (H,W,3)
    m = mask[0].to(torch.float32)
    pp = F.pad(pos.permute(2, 0, 1)[None], (1, 1, 1, 1), mode="replicate")[0].permute(1, 2, 0)
    mp = F.pad(m[None, None], (1, 1, 1, 1), mode="replicate")[0, 0].bool()

    def nb(dy, dx):
        return pp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W], mp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]

    c = pp[1:-1, 1:-1] * mp[1:-1, 1:-1, None]
    pu, mu = nb(-1, 0)
    pl, ml = nb(0, -1)
    pb, mb = nb(1, 0)
    pr, mr = nb(0, 1)
    u = (pu - c) * mu[..., None]
    l = (pl - c) * ml[..., None]
    b = (pb - c) * mb[..., None]
    r = (pr - c) * mr[..., None]
    n = (torch.linalg.cross(u, l) + torch.linalg.cross(r, u)
         + torch.linalg.cross(b, r) + torch.linalg.cross(l, b))
    n = F.normalize(n, dim=-1)
    return (n * mask[0][..., None]).permute(2, 0, 1)


# ---------------------------------------------------------------- one rendered view
def render_view(rasterize_fn, cam, fov, view, proj, render_mask, bg, attrs, hw,
                front_only=False, require_importance=False, weight_thres=0.03):
    """utils/operations.py:645-720 (render_cuda_core).  attrs = (means, harmonics(N,1,3),
    opacities(N,), confidences(N,), scales, rotations) already ACTIVATED.  rasterize_fn has the
    keyword surface of oracle.rasterizer_ref.rasterize."""
    means, harm, opac, conf, scales, rots = attrs
    tan = (0.5 * fov).tan()
    means2d = torch.zeros_like(means, requires_grad=means.requires_grad)
    rgb, normal, depth, opacity, confidence, importance, count, radii = rasterize_fn(
        means, means2d, opac[..., None], conf, harm[:, 0, :], scales, rots,
        image_height=hw[0], image_width=hw[1], tanfovx=float(tan[0]), tanfovy=float(tan[1]),
        bg=bg, viewmatrix=view, projmatrix=proj, scale_modifier=1.0, render_mask=render_mask,
        weight_thres=weight_thres, require_importance=require_importance, front_only=front_only)
    mask = opacity.detach() > 1e-2                                       # Q3: 1e-2 here
    normal = F.normalize(normal, dim=0) * mask
    d2n = depth2normal(depth, mask, fov)
    return rgb, depth, normal, opacity, d2n, confidence, importance, count, radii


def render_view_all(rasterize_fn, extrinsics, intrinsics, attrs, bg, near_far, hw,
                    render_masks=None, require_grad=False, require_importance=False,
                    front_only=False):
    """utils/operations.py:724-778 + 829-904 (GaussianRenderer.__init__ + render_view_all)."""
    B = extrinsics.shape[0]
    fovs, views, projs, campos = camera_setup(extrinsics, intrinsics, near_far)
    outs = []
    with torch.set_grad_enabled(require_grad):
        for i in range(B):
            rm = None if render_masks is None else render_masks[i]
            outs.append(render_view(rasterize_fn, campos[i], fovs[i], views[i], projs[i], rm, bg,
                                    attrs, hw, front_only, require_importance))
    st = lambda k: torch.stack([o[k] for o in outs])
    radii = torch.stack([o[8] for o in outs]).sum(0)
    return st(0), st(1), st(2), st(3), st(4), st(5), st(6), st(7), radii > 0


# ---------------------------------------------------------------- activations
def activate(means, scales_raw, rots_raw, opac_raw, harmonics, view_scores, view_supports,
             view_means, scale_factor=0.01, use_view_distribution=True):
    """mapping/gaussian_map.py:529-581 (get_attr): returns (means, harmonics, opacities,
    confidences, scales, rotations)."""
    rot = F.normalize(rots_raw)
    sc = torch.clamp(scale_factor * torch.exp(scales_raw), min=0, max=0.05)
    op = torch.sigmoid(opac_raw)
    if use_view_distribution:
        vv = view_means.norm(dim=-1)
        vv = torch.where(torch.isnan(vv), torch.ones_like(vv), vv)
        conf = torch.clamp(torch.exp(1 - vv) * view_scores, min=0, max=1)
    else:
        conf = torch.clamp(1 - 1 / torch.exp(view_supports), min=0, max=1)
    return means, harmonics, op, conf, sc, rot


# ---------------------------------------------------------------- losses
def central_diff(m):
    """mapping/utils.py:43-62 -- squared norms of the 4 one-sided differences, (B,4,H,W)."""
    dl = F.pad(m[..., :-1] - m[..., 1:], (0, 1))
    dr = F.pad(m[..., 1:] - m[..., :-1], (1, 0))
    du = F.pad(m[..., :-1, :] - m[..., 1:, :], (0, 0, 0, 1))
    dd = F.pad(m[..., 1:, :] - m[..., :-1, :], (0, 0, 1, 0))
    return (torch.stack([dl, dr, du, dd], 2) ** 2).sum(1)


def normal_tv_loss(normals, depths, mask, sigma=0.3):
    """mapping/utils.py:28-40."""
    nd = central_diff(normals)
    dd = central_diff(depths.detach())
    return torch.mean((dd <= 1e-4).float() * torch.exp(-nd / (2 * sigma ** 2)) * nd * mask)


def train_loss(rgb_p, depth_p, normal_p, opacity_p, d2n_p, rgb_gt, depth_gt):
    """mapping/gaussian_map.py:106-124.  Returns (total, per_frame_perf) where per_frame_perf is
    what track_performance (:132-139) writes.  Keeps quirk Q1: the (B,H,W) consistency map times
    the (B,1,H,W) integer mask broadcasts to (B,B,H,W)."""
    m_vis = opacity_p.detach() > 1e-3
    m_d = depth_gt > 0.0
    l_rgb = torch.abs((rgb_p - rgb_gt) * m_vis)
    l_d = torch.abs((depth_p - depth_gt) * m_d)
    perf = l_rgb.mean(dim=[1, 2, 3]).detach() + l_d.mean(dim=[1, 2, 3]).detach()
    tv = normal_tv_loss(normal_p, depth_p, m_d)
    cons = 1 - (normal_p * d2n_p).sum(1)
    cons = (cons * m_vis.long()).mean()
    total = l_rgb.mean() + 0.8 * l_d.mean() + 0.1 * cons + 0.1 * tv
    return total, perf


def cal_psnr(pred, gt):
    """mapping/utils.py:269-277."""
    mse = ((pred - gt) ** 2).mean().item()
    return -10 * math.log10(mse + 1e-8)


# ---------------------------------------------------------------- sampler
class WeightedSampler:
    """mapping/utils.py:190-228: newest `active_size` frames + up to batch-active older frames
    drawn without replacement with p ~ training_performance (np.random.choice).  NB the reference
    indexes random_ids_all with the drawn *values* (ids = random_ids_all[indices]); since
    random_ids_all == arange(k) that is the identity, restated as such."""

    def __init__(self, n_frames, batch_size=8, active_size=3):
        a = min(active_size, n_frames)
        ids = np.arange(n_frames)
        self.active_ids = ids[-a:]
        self.random_ids_all = ids[:-a]
        self.selected_num = min(len(self.random_ids_all), batch_size - a)

    def next_ids(self, weight):
        sel = self.active_ids.copy()
        if self.selected_num > 0:
            w = weight[self.random_ids_all]
            w = w / torch.sum(w)
            drawn = np.random.choice(self.random_ids_all, size=self.selected_num,
                                     p=w.cpu().numpy(), replace=False)
            sel = np.append(sel, drawn)
        return sel


# ---------------------------------------------------------------- Adam
LR = dict(mean=5e-4, scale=1e-2, rotation=5e-4, opacity=1e-2, harmonic=1e-4)  # incremental.yaml:27-32


def make_adam(means, scales, rots, opac, harm):
    """mapping/gaussian_map.py:259-292: five groups, eps=1e-15, fresh state."""
    return torch.optim.Adam(
        [dict(params=[means], lr=LR["mean"]), dict(params=[scales], lr=LR["scale"]),
         dict(params=[rots], lr=LR["rotation"]), dict(params=[opac], lr=LR["opacity"]),
         dict(params=[harm], lr=LR["harmonic"])], eps=1e-15)


def adam_step_ref(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-15):
    """torch.optim.Adam single-tensor update (no weight decay / amsgrad), written out."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


# ---------------------------------------------------------------- train loop
def train_iterations(state, frames, frame_id_batches, bg, near_far, hw, rasterize_fn=None,
                     scale_factor=0.01):
    """mapping/gaussian_map.py:66-127 with the keyframe ids per iteration given explicitly
    (so the RNG of the sampler is outside).  `state` = dict of raw tensors
    (means, scales, rotations, opacities, harmonics(N,1,3), view_scores, view_supports,
    view_means); updated in place.  Returns list of (loss, perf) per iteration."""
    rasterize_fn = rasterize_fn or rr.rasterize
    names = ["means", "scales", "rotations", "opacities", "harmonics"]
    params = [torch.nn.Parameter(state[k].clone()) for k in names]
    opt = make_adam(*params)
    log = []
    for ids in frame_id_batches:
        rgb_gt = torch.stack([frames[i]["rgb"] for i in ids])
        d_gt = torch.stack([frames[i]["depth"] for i in ids])
        ext = torch.stack([frames[i]["extrinsic"] for i in ids])
        intr = torch.stack([frames[i]["intrinsic"] for i in ids])
        attrs = activate(params[0], params[1], params[2], params[3], params[4],
                         state["view_scores"], state["view_supports"], state["view_means"],
                         scale_factor)
        rgb, depth, normal, opacity, d2n, *_ = render_view_all(
            rasterize_fn, ext, intr, attrs, bg, near_far, hw, require_grad=True)
        loss, perf = train_loss(rgb, depth, normal, opacity, d2n, rgb_gt, d_gt)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        log.append((float(loss.detach()), perf))
    for k, p in zip(names, params):
        state[k] = p.detach()
    return log


# ---------------------------------------------------------------- post-processing + prune
def view_stats_update(state, upd, cam_pos, depth_max, use_view_distribution=True):
    """mapping/gaussian_map.py:195-227: confidence bookkeeping for the Gaussians counted in the newest
    keyframe (`upd` (N,) bool): supports += 1; running mean of the unit view direction; score +=
    (1 - clamp(dist/depth_max)) * clamp(normal . dir).  Updates `state` in place."""
    state["view_supports"] = state["view_supports"] + upd.float()
    if not use_view_distribution:
        return
    normals = F.normalize(rr.quat_to_rotmat(F.normalize(state["rotations"]))[:, :3, 2])
    vdir = cam_pos[None] - state["means"]
    dist = torch.linalg.norm(vdir, dim=1)
    vdir = vdir / dist[:, None]
    vm = state["view_means"].clone()
    vm[upd] += (vdir[upd] - vm[upd]) / state["view_supports"][upd][:, None]
    state["view_means"] = vm
    cos = torch.clamp((normals * vdir).sum(1), 0, 1)
    dfac = torch.clamp(dist / depth_max, 0, 1)
    vs = state["view_scores"].clone()
    vs[upd] += ((1 - dfac) * cos)[upd]
    state["view_scores"] = vs


def post_process(state, frames, bg, near_far, hw, prune_interval=5, rasterize_fn=None,
                 scale_factor=0.01, use_view_distribution=True):
    """mapping/gaussian_map.py:141-246.  Confidence bookkeeping from the newest keyframe's
    visibility counts; every `prune_interval`-th keyframe all T frames are rendered and Gaussians
    never counted, or with activated opacity < 0.1, are removed.  `state` is updated in place;
    returns the boolean keep-mask (all True when no prune ran)."""
    rasterize_fn = rasterize_fn or rr.rasterize
    T = len(frames)
    prune = (T % prune_interval) == 0
    ids = list(range(T)) if prune else [T - 1]
    ext = torch.stack([frames[i]["extrinsic"] for i in ids])
    intr = torch.stack([frames[i]["intrinsic"] for i in ids])
    d_gt = torch.stack([frames[i]["depth"] for i in ids])
    attrs = activate(state["means"], state["scales"], state["rotations"], state["opacities"],
                     state["harmonics"], state["view_scores"], state["view_supports"],
                     state["view_means"], scale_factor, use_view_distribution)
    counts = render_view_all(rasterize_fn, ext, intr, attrs, bg, near_far, hw,
                             render_masks=(d_gt > 0.0).float(), require_importance=True,
                             front_only=True)[7]
    view_stats_update(state, counts[-1] >= 1.0, ext[-1, :3, 3], frames[-1]["depth_range"][1], use_view_distribution)
    keep = torch.ones(state["means"].shape[0], dtype=torch.bool)
    if prune:
        never_seen = ~(counts.sum(0) >= 1.0)
        keep = ~(never_seen | (torch.sigmoid(state["opacities"]) < 0.1))
        for k in ["means", "scales", "rotations", "opacities", "harmonics", "view_scores",
                  "view_supports", "view_means"]:
            state[k] = state[k][keep]
    return keep


# ---------------------------------------------------------------- spawn (add_gaussians)
def smooth_depth(depth, tolerance=0.5):
    """utils/operations.py:161-169 -- OpenCV bilateral filter (d=15, sigmaColor=tolerance,
    sigmaSpace=20) of the zero-filled depth, invalid (< 0) pixels back to -1.  (1,H,W) -> (1,H,W)."""
    import cv2
    d = depth.squeeze(0).cpu().numpy()
    bad = d < 0.0
    filt = cv2.bilateralFilter(np.where(bad, 0.0, d).astype(np.float32), 15, tolerance, 20)
    filt[bad] = -1.0
    return torch.tensor(filt).unsqueeze(0)


def pixel_world_rays(H, W, extrinsic, intrinsic):
    """utils/operations.py:372-392 (pixel centres (x+.5)/W, (y+.5)/H), :464-478 (K^-1 [x,y,1]) and
    :544-569 (rotate into the world, origin = camera position): (H*W,3) origins, directions
    (un-normalised, camera z = 1), row-major pixel order."""
    ys = (torch.arange(H, dtype=torch.float32) + 0.5) / H
    xs = (torch.arange(W, dtype=torch.float32) + 0.5) / W
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([xx, yy, torch.ones_like(xx)], -1).reshape(-1, 3)
    d_cam = pix @ torch.linalg.inv(intrinsic).t()
    d_world = d_cam @ extrinsic[:3, :3].t()
    return extrinsic[:3, 3].expand_as(d_world), d_world


def normal_to_quaternion(n):
    """utils/operations.py:481-500 (normal2rotation) + :526-541 (rotmat2quaternion): frame with z = n,
    x = the world x axis (y axis when |n_x| > 0.99) made orthogonal to n; quaternion (r,x,y,z) from
    the trace formula with the +1e-6 the reference adds, normalised."""
    z = n / n.norm(dim=1, keepdim=True)
    ref = torch.zeros_like(z)
    ref[:, 0] = 1.0
    ref[z[:, 0].abs() > 0.99] = torch.tensor([0.0, 1.0, 0.0])
    x = ref - (ref * z).sum(1, keepdim=True) * z
    x = x / x.norm(dim=1, keepdim=True)
    y = torch.linalg.cross(z, x, dim=1)
    y = y / y.norm(dim=1, keepdim=True)
    R = torch.stack([x, y, z], dim=-1)
    r = torch.sqrt(1 + (R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2] + 1e-6)) / 2
    q = torch.stack([r, (R[:, 2, 1] - R[:, 1, 2]) / (4 * r), (R[:, 0, 2] - R[:, 2, 0]) / (4 * r),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * r)], -1)
    return F.normalize(q, dim=-1)


def spawn_mask(rgb_gt, depth_gt, pred, error_thres=0.25):
    """mapping/gaussian_map.py:470-489 (cal_mask) for one frame: spawn where the map renders the wrong
    colour (MSE over channels > error_thres), is not opaque (< 0.5) or lies more than 5 % BEHIND the
    sensor depth.  pred = None (no map yet) selects every pixel.  Returns (H*W,) bool."""
    H, W = rgb_gt.shape[1:]
    if pred is None:
        return torch.ones(H * W, dtype=torch.bool)
    err = ((rgb_gt - pred["rgb"]) ** 2).mean(0)
    m = err > error_thres
    m = m | (pred["opacity"] < 0.5)
    m = m | ((depth_gt[0] - pred["depth"]) < -0.05 * depth_gt[0])
    return m.reshape(-1)


def spawn_candidates(frame, pred=None, error_thres=0.25, depth_smooth=None):
    """mapping/gaussian_map.py:294-400 up to (not including) the random voxel filter: the per-pixel
    candidate Gaussians of one RGB-D keyframe.  Returns dict(select (H*W,) bool, means (H*W,3),
    rotations (H*W,4), colors (H*W,3)); the new Gaussians are rows `select` in pixel order with raw
    scales (0,0,-1e10), raw opacity 0, and zero view statistics (:369-381)."""
    rgb, depth = frame["rgb"], frame["depth"]
    _, H, W = rgb.shape
    if depth_smooth is None:
        depth_smooth = smooth_depth(depth)
    valid = (depth > 0.0).reshape(-1)
    origins, directions = pixel_world_rays(H, W, frame["extrinsic"], frame["intrinsic"])
    pcd = origins + directions * depth.reshape(-1, 1)
    # normals from the smoothed depth with a hard-wired 60 x 60 degree fov (:316-322)
    n_cam = depth2normal(depth_smooth, valid.view(1, H, W), (np.pi / 3, np.pi / 3)).permute(1, 2, 0).reshape(-1, 3)
    valid = valid & ((n_cam ** 2).sum(-1) > 0.0)
    n_world = n_cam @ frame["extrinsic"][:3, :3].t()
    normals = torch.zeros(H * W, 3)
    normals[:, 2] = 1.0
    normals[valid] = n_world[valid]
    cos = (F.normalize(directions, dim=1) * normals).sum(-1)              # back-facing normals out (:330-335)
    valid = valid & (cos < -0.01)
    q = normal_to_quaternion(normals)
    valid = valid & ~torch.any(q.isnan(), dim=1)
    select = spawn_mask(rgb, depth, pred, error_thres) & valid
    return dict(select=select, means=pcd, rotations=q, colors=rgb.permute(1, 2, 0).reshape(-1, 3))


def voxel_ids(points, voxel_size=0.02):
    """utils/operations.py:603-606: integer voxel coordinates floor(p / voxel_size), (n,3) int64."""
    return torch.floor(points / voxel_size).long()


def voxel_filter_is_valid(points, selected, voxel_size=0.02):
    """What voxel_downsample (utils/operations.py:603-625) guarantees whatever its random draw:
    `selected` is sorted, holds exactly one point of every occupied voxel and nothing else."""
    vox = voxel_ids(points, voxel_size)
    uniq = torch.unique(vox, dim=0)
    sel = torch.as_tensor(selected).long()
    if sel.numel() != uniq.shape[0] or not bool((sel[1:] > sel[:-1]).all()):
        return False
    return torch.unique(vox[sel], dim=0).shape[0] == uniq.shape[0]


# ---------------------------------------------------------------- planner utilities (section 8 f1)
def voxel_visible_mask(voxel_centers, extrinsic, intrinsic, depth):
    """mapping/voxel_map.py:226-278 (cal_visible_mask): voxel centres in front of the camera that
    project inside the image (normalised K; x*w, y*h truncated to the pixel index) and lie in front of
    the depth stored there.  depth (h,w); returns (M,) bool."""
    h, w = depth.shape
    hom = torch.cat([voxel_centers, torch.ones(voxel_centers.shape[0], 1)], -1)
    cam = (torch.linalg.inv(extrinsic) @ hom.t()).t()[:, :3]
    z = cam[:, 2]
    img = (intrinsic @ cam.t()).t()
    xy = img[:, :2] / img[:, 2:3]
    x, y = xy[:, 0] * w, xy[:, 1] * h
    inside = (x >= 0) & (x < w) & (y >= 0) & (y < h)
    dval = torch.full((voxel_centers.shape[0],), -1.0)
    dval[inside] = depth[y[inside].long(), x[inside].long()]
    return (z > 0) & inside & (dval > z)


def view_utilities(depth, confidence, voxel_centers, unexplored, extrinsics, intrinsics, depth_range,
                   valid_mask=None):
    """planning/confidence.py:69-101 (and planning/exploration.py:62-86, which is the exploration
    half alone) for V rendered candidate views: depth, confidence (V,h,w).  Returns
    (explore (V,), exploit (V,)): the fraction of all voxels that are visible AND unexplored, and the
    mean distance-weighted uncertainty (1 - confidence)."""
    V, h, w = depth.shape
    lo, hi = float(depth_range[0]), float(depth_range[1])
    explore, exploit = torch.zeros(V), torch.zeros(V)
    for i in range(V):
        valid = torch.ones(h, w, dtype=torch.bool) if valid_mask is None else valid_mask[i]
        dv = depth[i].clone()
        dv[dv < 0.001] = 10000.0                                           # nothing rendered: free up to the range
        dv = dv.clamp(lo, hi)
        dv[~valid] = -1.0
        vis = voxel_visible_mask(voxel_centers, extrinsics[i], intrinsics[i], dv)
        explore[i] = (vis & unexplored).sum() / voxel_centers.shape[0]
        conf = confidence[i].clone()
        conf[depth[i] > hi] = 1.0
        conf[~valid] = 1.0
        ds = depth[i].clone()
        ds[ds < 0.001] = hi * 0.5
        exploit[i] = ((1 - conf) * ds / hi).mean()
    explore[torch.isnan(explore)] = 0.0
    exploit[torch.isnan(exploit)] = 0.0
    return explore, exploit


def low_confidence_voxels(state, bbox_min, size, dim, min_gaussian_per_voxel=5, confidence_thres=0.3,
                          opacity_thres=0.7, scale_factor=0.01):
    """mapping/voxel_map.py:70-113 (the Gaussian half of VoxelMap.update_utility): per voxel the number
    of opaque (> opacity_thres) low-confidence (< confidence_thres) Gaussians whose mean falls in it,
    update_mask = count > min_gaussian_per_voxel, voxel_normal = normalised mean of their normals
    (0 elsewhere).  Returns (count (M,) int64, voxel_normal (M,3), update_mask (M,) bool)."""
    means, _, opac, conf, _, rot = activate(state["means"], state["scales"], state["rotations"], state["opacities"],
                                            state["harmonics"], state["view_scores"], state["view_supports"],
                                            state["view_means"], scale_factor)
    normals = F.normalize(rr.quat_to_rotmat(rot)[:, :3, 2])
    dim = torch.as_tensor(dim).int()
    idx = torch.floor((means - bbox_min) / size).int()
    ok = torch.all(idx >= 0, dim=1) & torch.all(idx < dim, dim=1) & (conf < confidence_thres) & (opac > opacity_thres)
    idx, normals = idx[ok].long(), normals[ok]
    lin = idx[:, 0] * int(dim[1] * dim[2]) + idx[:, 1] * int(dim[2]) + idx[:, 2]
    M = int(torch.prod(dim))
    count = torch.zeros(M, dtype=torch.int64).scatter_add(0, lin, torch.ones_like(lin))
    nsum = torch.zeros(M, 3).scatter_add(0, lin[:, None].expand(-1, 3), normals)
    upd = count > min_gaussian_per_voxel
    vn = torch.zeros(M, 3)
    vn[upd] = F.normalize(nsum[upd] / count[upd, None], dim=-1)
    return count, vn, upd
