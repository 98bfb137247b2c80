"""CPU oracle of the native half of the hot path (TEST INFRASTRUCTURE ONLY).

    *** PARITY UNPINNED ***  The arithmetic this file restates lives in the third-party CUDA
    extension `diff_gaussian_rasterization_2d` (git+https://github.com/liren-jin/
    diff-gaussian-rasterization_2d, pinned version: NONE -- /root/reference/envs/requirements.txt:15).
    Its source is not under /root/reference and the reference ships no tests / golden vectors for
    it.  The semantics below are therefore OUR specification,
    constrained by the reference's only call site (/root/reference/utils/operations.py:682-720) and
    the public 3DGS -> GaussianSurfels lineage.  Only tests/, __graft_entry__.smoke() and bench.py's
    cpu_baseline / --impl reference leg may import this module; the product path never does.

Pure PyTorch, differentiable through autograd (this is the gradient truth for the CUDA backward),
dtype-generic (float32 / float64).  One view per call.

Boundary restated (operations.py:682-713):
    settings: image_height, image_width, tanfovx, tanfovy, bg(3|4), scale_modifier, viewmatrix
              (= (w2c)^T, row-vector convention), projmatrix (= viewmatrix @ P^T), sh_degree(0),
              campos, prefiltered, render_mask ((0,) or (1,H,W)), weight_thres, debug,
              config(5) = [1,1,1,require_importance,front_only]
    call    : means3D(N,3) means2D(N,3) opacities(N,1) confidences(N,) shs=None
              colors_precomp(N,3) scales(N,3) rotations(N,4) cov3D_precomp=None
    returns : rgb(3,H,W) normal(3,H,W) depth(1,H,W) opacity(1,H,W) confidence(1,H,W)
              importance(N,) f32  count(N,) i32  radii(N,) i32
"""
import math
import torch

TILE = 16                 # tile edge in pixels (part of the semantics: rect culling is per tile)
NEAR_CULL = 0.2           # view-space z cull (3DGS lineage)
LOWPASS = 0.3             # screen-space dilation added to the 2D covariance diagonal
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0
T_EPS = 1e-4              # stop compositing when transmittance would fall below this
SLOPE_COS_MIN = 0.1       # clamp on n.(t/t_z) used by the per-pixel depth plane


def quat_to_rotmat(q):
    """(r,x,y,z) -> R, same element order as /root/reference/utils/operations.py:261-278."""
    r, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y),
        ],
        -1,
    ).reshape(q.shape[:-1] + (3, 3))
    return R


def project(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
            scale_modifier=1.0, front_only=False):
    """Per-Gaussian projection (kernel K1).  Returns a dict of per-Gaussian screen-space
    quantities plus the integer radius / tile rect.  Differentiable where meaningful."""
    dt = means3D.dtype
    N = means3D.shape[0]
    V = viewmatrix.to(dt)
    M = projmatrix.to(dt)
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)

    t = means3D @ V[:3, :3] + V[3, :3]                      # view-space centre
    p_hom = means3D @ M[:3, :] + M[3, :]
    inv_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc_x = p_hom[:, 0] * inv_w
    ndc_y = p_hom[:, 1] * inv_w
    px = ((ndc_x + 1.0) * W - 1.0) * 0.5                    # pixel centres at integers
    py = ((ndc_y + 1.0) * H - 1.0) * 0.5

    R = quat_to_rotmat(rotations)                           # rotations used as given (caller normalises)
    s = scales * scale_modifier
    Sigma = (R * (s * s)[:, None, :]) @ R.transpose(1, 2)   # R diag(s^2) R^T ; third scale may be 0

    Wr = V[:3, :3].t()                                      # world->view rotation (column convention)
    tz = t[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack(
        [fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], -1
    ).reshape(N, 2, 3)
    Tm = J @ Wr                                             # (N,2,3)
    cov2 = Tm @ Sigma @ Tm.transpose(1, 2)
    a = cov2[:, 0, 0] + LOWPASS
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + LOWPASS
    det = a * c - b * b
    det_safe = torch.where(det == 0, torch.ones_like(det), det)
    conic = torch.stack([c / det_safe, -b / det_safe, a / det_safe], -1)
    mid = 0.5 * (a + c)
    lam1 = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam1.detach())).to(torch.int64)

    # view-space normal = third column of R, flipped toward the camera
    n_w = R[:, :, 2]
    n_v = n_w @ Wr.t()
    cosv = (n_v * t).sum(-1)
    back_facing = cosv >= 0
    n_v = torch.where((cosv > 0)[:, None], -n_v, n_v)
    c0 = (n_v * t).sum(-1)                                  # <= 0 after the flip
    Dc = torch.clamp(c0 / tz, max=-SLOPE_COS_MIN)
    slope_x = -tz * n_v[:, 0] / (Dc * fx)                   # d depth / d pixel-x  (first order)
    slope_y = -tz * n_v[:, 1] / (Dc * fy)

    # culling + tile rect
    valid = (tz.detach() > NEAR_CULL) & (det.detach() != 0)
    if front_only:
        valid = valid & ~back_facing
    tiles_x = (W + TILE - 1) // TILE
    tiles_y = (H + TILE - 1) // TILE
    pxd, pyd = px.detach(), py.detach()
    rf = radius.to(dt)
    # (int) cast truncates toward zero, like the C cast in the kernel
    rminx = torch.clamp(torch.trunc((pxd - rf) / TILE).to(torch.int64), 0, tiles_x)
    rminy = torch.clamp(torch.trunc((pyd - rf) / TILE).to(torch.int64), 0, tiles_y)
    rmaxx = torch.clamp(torch.trunc((pxd + rf + TILE - 1) / TILE).to(torch.int64), 0, tiles_x)
    rmaxy = torch.clamp(torch.trunc((pyd + rf + TILE - 1) / TILE).to(torch.int64), 0, tiles_y)
    ntiles = (rmaxx - rminx) * (rmaxy - rminy)
    valid = valid & (ntiles > 0)
    radius = torch.where(valid, radius, torch.zeros_like(radius))
    return dict(px=px, py=py, conic=conic, depth=tz, slope_x=slope_x, slope_y=slope_y,
                normal=n_v, radius=radius, valid=valid,
                rect=(rminx, rminy, rmaxx, rmaxy), tiles=(tiles_x, tiles_y))


def rasterize(means3D, means2D, opacities, confidences, colors, scales, rotations, *,
              image_height, image_width, tanfovx, tanfovy, bg, viewmatrix, projmatrix,
              scale_modifier=1.0, render_mask=None, weight_thres=0.03,
              require_importance=False, front_only=False):
    """Full native forward of one view.  `means2D` only receives a gradient: the screen-space
    position used is px + means2D[:,0], py + means2D[:,1] with means2D == 0 (so its .grad is
    dL/d(pixel-space mean), operations.py:676-680)."""
    dt = means3D.dtype
    dev = means3D.device
    H, W = int(image_height), int(image_width)
    N = means3D.shape[0]
    P = project(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                scale_modifier, front_only)
    px = P["px"] + (means2D[:, 0] if means2D is not None else 0)
    py = P["py"] + (means2D[:, 1] if means2D is not None else 0)
    tiles_x, tiles_y = P["tiles"]
    rminx, rminy, rmaxx, rmaxy = P["rect"]
    opac = opacities.reshape(N)
    bg3 = bg.to(dt)[:3]

    out_rgb = torch.zeros(3, H, W, dtype=dt, device=dev) + bg3[:, None, None]
    out_n = torch.zeros(3, H, W, dtype=dt, device=dev)
    out_d = torch.zeros(1, H, W, dtype=dt, device=dev)
    out_a = torch.zeros(1, H, W, dtype=dt, device=dev)
    out_c = torch.zeros(1, H, W, dtype=dt, device=dev)
    importance = torch.zeros(N, dtype=torch.float32, device=dev)
    count = torch.zeros(N, dtype=torch.int32, device=dev)

    # instance list: (tile, depth-as-float32, id) sorted; ties keep index order (stable radix sort)
    vid = torch.nonzero(P["valid"]).flatten()
    if vid.numel() > 0:
        nx = (rmaxx - rminx)[vid]
        ny = (rmaxy - rminy)[vid]
        cnt = nx * ny
        inst_g = torch.repeat_interleave(vid, cnt)
        start = torch.cumsum(cnt, 0) - cnt
        local = torch.arange(inst_g.numel(), device=dev) - torch.repeat_interleave(start, cnt)
        nx_i = torch.repeat_interleave(nx, cnt)
        tx_i = rminx[inst_g] + local % nx_i
        ty_i = rminy[inst_g] + local // nx_i
        tile_i = ty_i * tiles_x + tx_i
        depth32 = P["depth"].detach().to(torch.float32)[inst_g]
        order = torch.argsort(inst_g, stable=True)
        order = order[torch.argsort(depth32[order], stable=True)]
        order = order[torch.argsort(tile_i[order], stable=True)]
        inst_g = inst_g[order]
        tile_i = tile_i[order]
        tile_ids, tile_cnt = torch.unique_consecutive(tile_i, return_counts=True)
        tile_start = torch.cumsum(tile_cnt, 0) - tile_cnt
    else:
        tile_ids = torch.zeros(0, dtype=torch.int64)
        tile_cnt = tile_start = tile_ids

    mask_flat = None
    if render_mask is not None and render_mask.numel() > 0:
        mask_flat = render_mask.reshape(H, W)

    for k in range(tile_ids.numel()):
        tile = int(tile_ids[k])
        g = inst_g[int(tile_start[k]): int(tile_start[k]) + int(tile_cnt[k])]
        ty0, tx0 = (tile // tiles_x) * TILE, (tile % tiles_x) * TILE
        ys = torch.arange(ty0, min(ty0 + TILE, H), device=dev)
        xs = torch.arange(tx0, min(tx0 + TILE, W), device=dev)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pxf = xx.reshape(-1, 1).to(dt)
        pyf = yy.reshape(-1, 1).to(dt)
        dx = px[g][None, :] - pxf                           # (npix, n)
        dy = py[g][None, :] - pyf
        con = P["conic"][g]
        power = -0.5 * (con[:, 0] * dx * dx + con[:, 2] * dy * dy) - con[:, 1] * dx * dy
        alpha = torch.clamp(opac[g][None, :] * torch.exp(power), max=ALPHA_MAX)
        skip = (power.detach() > 0) | (alpha.detach() < ALPHA_MIN)
        alpha = torch.where(skip, torch.zeros_like(alpha), alpha)
        one_m = 1.0 - alpha
        T_incl = torch.cumprod(one_m, dim=1)                # T after applying j
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], 1)
        # termination: first non-skipped j with T_incl < T_EPS is NOT applied, nor anything after
        dead = (~skip) & (T_incl.detach() < T_EPS)
        alive = torch.cumsum(dead.to(torch.int32), 1) == 0
        w = alpha * T_excl * alive.to(dt)                   # blending weights
        contrib = alive & ~skip
        # final transmittance = T_incl at the last alive column (1 if none)
        n_alive = alive.to(torch.int64).sum(1)
        T_pad = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl], 1)
        T_final = torch.gather(T_pad, 1, n_alive[:, None])[:, 0]

        depth_pix = (P["depth"][g][None, :] - P["slope_x"][g][None, :] * dx
                     - P["slope_y"][g][None, :] * dy)       # z_c + s.(pix - centre)
        C = w @ colors[g]                                   # (npix,3)
        Nn = w @ P["normal"][g]
        D = (w * depth_pix).sum(1)
        Cf = w @ confidences[g].to(dt)
        A = 1.0 - T_final
        depth_out = torch.where(A > 0, D / torch.where(A > 0, A, torch.ones_like(A)),
                                torch.zeros_like(D))
        hh, ww = ys.numel(), xs.numel()
        sl = (slice(None), slice(ty0, ty0 + hh), slice(tx0, tx0 + ww))
        out_rgb[sl] = (C + T_final[:, None] * bg3[None, :]).t().reshape(3, hh, ww)
        out_n[sl] = Nn.t().reshape(3, hh, ww)
        out_d[sl] = depth_out.reshape(1, hh, ww)
        out_a[sl] = A.reshape(1, hh, ww)
        out_c[sl] = Cf.reshape(1, hh, ww)

        if require_importance:
            hit = contrib & (w.detach() > weight_thres)
            if mask_flat is not None:
                m = mask_flat[ty0:ty0 + hh, tx0:tx0 + ww].reshape(-1, 1) == 1
                hit = hit & m
            count.index_add_(0, g, hit.sum(0).to(torch.int32))
            importance.index_add_(0, g, (w.detach() * hit).sum(0).to(torch.float32))

    radii = P["radius"].to(torch.int32)
    return out_rgb, out_n, out_d, out_a, out_c, importance, count, radii


def num_instances(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                  scale_modifier=1.0, front_only=False):
    """(V visible, I tile instances) for one view -- the sizes the roofline formulas use."""
    P = project(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                scale_modifier, front_only)
    rminx, rminy, rmaxx, rmaxy = P["rect"]
    n = ((rmaxx - rminx) * (rmaxy - rminy))[P["valid"]]
    return int(P["valid"].sum()), int(n.sum())


# ----------------------------------------------------------------------------------------------
# Boundary-shaped wrappers so the reference's render_cuda_core body runs on the oracle unchanged.
class GaussianRasterizationSettings:
    def __init__(self, image_height, image_width, tanfovx, tanfovy, bg, scale_modifier,
                 viewmatrix, projmatrix, sh_degree, campos, prefiltered, render_mask,
                 weight_thres, debug, config):
        self.image_height, self.image_width = image_height, image_width
        self.tanfovx, self.tanfovy = tanfovx, tanfovy
        self.bg, self.scale_modifier = bg, scale_modifier
        self.viewmatrix, self.projmatrix = viewmatrix, projmatrix
        self.sh_degree, self.campos, self.prefiltered = sh_degree, campos, prefiltered
        self.render_mask, self.weight_thres, self.debug, self.config = (
            render_mask, weight_thres, debug, config)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, confidences, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        s = self.raster_settings
        if shs is not None or colors_precomp is None:
            raise ValueError("oracle supports colors_precomp only (sh_degree=0 path)")
        if cov3D_precomp is not None or scales is None or rotations is None:
            raise ValueError("oracle supports scales/rotations only")
        cfg = [float(v) for v in s.config.detach().cpu().tolist()]
        return rasterize(
            means3D, means2D, opacities, confidences, colors_precomp, scales, rotations,
            image_height=s.image_height, image_width=s.image_width, tanfovx=s.tanfovx,
            tanfovy=s.tanfovy, bg=s.bg, viewmatrix=s.viewmatrix, projmatrix=s.projmatrix,
            scale_modifier=s.scale_modifier, render_mask=s.render_mask,
            weight_thres=s.weight_thres, require_importance=cfg[3] > 0, front_only=cfg[4] > 0)
