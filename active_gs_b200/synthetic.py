"""Deterministic synthetic 'Replica-shaped' scenes, cameras and dataframes.

No Replica / habitat-sim in this environment, so every workload is generated: Gaussian surfels on
the inner faces of a room box plus a few interior boxes, OpenCV c2w cameras inside the room, and
the dataframe dict layout the reference simulator emits
(/root/reference/simulator/habitat_simulator.py:105-136): rgb (3,H,W) in [0,1], depth (1,H,W)
metres with -1 = out of range, extrinsic (4,4) c2w OpenCV, intrinsic (3,3) normalised by W/H,
depth_range (2,).  Everything here is CPU/numpy so fixtures are reproducible bit-for-bit.
"""
import math
import numpy as np
import torch

ROOMS = {  # config index (BASELINE.json) -> (box xyz metres, H, W, N)
    2: ((6.0, 4.5, 2.7), 480, 640, 200_000),
    3: ((8.0, 6.0, 2.8), 720, 1280, 500_000),
    4: ((6.0, 4.5, 2.7), 480, 640, 200_000),
    5: ((12.0, 9.0, 3.0), 1080, 1920, 1_000_000),
}


def normal2rotation(n):
    """Unit quaternion (r,x,y,z) whose rotation's third column is n.  Restates
    /root/reference/utils/operations.py:481-541 (normal2rotation + rotmat2quaternion)."""
    z = n / n.norm(dim=1, keepdim=True)
    ref = torch.zeros_like(z)
    ref[:, 0] = 1.0
    par = z[:, 0].abs() > 0.99
    ref[par] = torch.tensor([0.0, 1.0, 0.0], dtype=z.dtype)
    x = ref - (ref * z).sum(1, keepdim=True) * z
    x = x / x.norm(dim=1, keepdim=True)
    y = torch.linalg.cross(z, x)
    y = y / y.norm(dim=1, keepdim=True)
    R = torch.stack([x, y, z], -1)
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2] + 1e-6
    r = torch.sqrt(1 + tr) / 2
    q = torch.stack([r, (R[:, 2, 1] - R[:, 1, 2]) / (4 * r), (R[:, 0, 2] - R[:, 2, 0]) / (4 * r),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * r)], -1)
    return torch.nn.functional.normalize(q, dim=-1)


def normalised_intrinsic(H, W, hfov_deg, vfov_deg=None):
    """/root/reference/simulator/utils.py:13-30 with normalize=True.  vfov defaults to fx == fy."""
    fx = (W / 2) / math.tan(math.radians(hfov_deg) / 2)
    fy = fx if vfov_deg is None else (H / 2) / math.tan(math.radians(vfov_deg) / 2)
    return torch.tensor([[fx / W, 0, 0.5], [0, fy / H, 0.5], [0, 0, 1]], dtype=torch.float32)


def look_c2w(pos, yaw, pitch):
    """OpenCV camera-to-world (x right, y down, z forward) in a z-up world."""
    f = np.array([math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), math.sin(pitch)])
    d0 = np.array([0.0, 0.0, -1.0])
    r = np.cross(d0, f)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = r, d, f, pos
    return torch.tensor(M, dtype=torch.float32)


def _box_surface(rng, n, lo, hi, inward):
    """n points on the 6 faces of [lo,hi] (area-weighted) with face normals (inward or outward)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    ext = hi - lo
    areas = np.array([ext[1] * ext[2]] * 2 + [ext[0] * ext[2]] * 2 + [ext[0] * ext[1]] * 2)
    face = rng.choice(6, size=n, p=areas / areas.sum())
    p = lo + rng.random((n, 3)) * ext
    nrm = np.zeros((n, 3))
    ax, side = face // 2, face % 2
    idx = np.arange(n)
    p[idx, ax] = np.where(side == 0, lo[ax], hi[ax])
    sign = np.where(side == 0, 1.0, -1.0) * (1.0 if inward else -1.0)
    nrm[idx, ax] = sign
    return p, nrm, face


def make_room_scene(N, box=(6.0, 4.5, 2.7), seed=1002, furniture=12):
    """Raw (pre-activation) GaussianMap state: surfels on room walls (80 %) + interior boxes."""
    rng = np.random.default_rng(seed)
    n_f = int(0.2 * N) if furniture > 0 else 0
    n_w = N - n_f
    pts, nrm, face = _box_surface(rng, n_w, (0, 0, 0), box, inward=True)
    base = rng.random((6 + furniture, 3))
    col = base[face]
    if n_f:
        per = np.full(furniture, n_f // furniture)
        per[: n_f - per.sum()] += 1
        for k, m in enumerate(per):
            size = rng.uniform(0.3, 1.2, 3)
            size[2] = rng.uniform(0.3, 1.0)
            lo = np.array([rng.uniform(0.3, box[0] - 0.3 - size[0]),
                           rng.uniform(0.3, box[1] - 0.3 - size[1]), 0.0])
            p, nn, _ = _box_surface(rng, m, lo, lo + size, inward=False)
            pts = np.concatenate([pts, p])
            nrm = np.concatenate([nrm, nn])
            col = np.concatenate([col, np.tile(base[6 + k], (m, 1))])
    nrm = nrm + rng.normal(0, 0.05, nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    col = np.clip(col + rng.normal(0, 0.05, col.shape), 0, 1)
    scales = np.stack([rng.uniform(-0.5, 0.7, N), rng.uniform(-0.5, 0.7, N), np.full(N, -1e10)], 1)
    vm = rng.normal(size=(N, 3))
    vm = vm / np.linalg.norm(vm, axis=1, keepdims=True) * rng.random((N, 1))
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    return dict(
        means=f32(pts), scales=f32(scales), rotations=normal2rotation(f32(nrm)),
        opacities=f32(rng.normal(2.0, 1.0, N)), harmonics=f32(col).reshape(N, 1, 3),
        view_scores=f32(rng.uniform(0, 2, N)), view_supports=f32(rng.poisson(3, N)),
        view_means=f32(vm))


def make_cameras(B, box=(6.0, 4.5, 2.7), H=480, W=640, hfov=60.0, seed=2002):
    rng = np.random.default_rng(seed)
    ext = []
    for _ in range(B):
        pos = np.array([rng.uniform(0.25, 0.75) * box[0], rng.uniform(0.25, 0.75) * box[1],
                        rng.uniform(1.2, 1.6)])
        ext.append(look_c2w(pos, rng.uniform(0, 2 * math.pi), math.radians(rng.normal(0, 10))))
    K = normalised_intrinsic(H, W, hfov)
    return torch.stack(ext), K[None].repeat(B, 1, 1)


def perturb_state(state, seed=3002):
    """Training start = generating scene perturbed: means +N(0,5mm), colour
    +N(0,0.1), opacity logit -1."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in state.items()}
    out["means"] += 0.005 * torch.randn(out["means"].shape, generator=g)
    out["harmonics"] += 0.1 * torch.randn(out["harmonics"].shape, generator=g)
    out["opacities"] -= 1.0
    return out


def noisy_depth(depth, seed=4002):
    """Sensor model of the synthetic GT: x(1+N(0,0.01)) (habitat.yaml:13) and 2 % pixels = -1."""
    g = torch.Generator().manual_seed(seed)
    d = depth * (1 + 0.01 * torch.randn(depth.shape, generator=g))
    drop = torch.rand(depth.shape, generator=g) < 0.02
    return torch.where(drop | (depth <= 0), torch.full_like(d, -1.0), d)


def make_c1_scene(seed=1001, N=1000, H=64, W=64):
    """BASELINE config 1: 1k random Gaussians in a 2 m cube 1.5-3.5 m in front of one camera,
    64x64, fov 60x60."""
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), rng.uniform(1.5, 3.5, N)], 1)
    nrm = -pts / np.linalg.norm(pts, axis=1, keepdims=True) + rng.normal(0, 0.3, (N, 3))
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    # larger disks than the room scenes so 1k surfels cover a 64x64 image
    scales = np.stack([rng.uniform(0.8, 1.6, N), rng.uniform(0.8, 1.6, N), np.full(N, -1e10)], 1)
    vm = rng.normal(size=(N, 3))
    vm = vm / np.linalg.norm(vm, axis=1, keepdims=True) * rng.random((N, 1))
    state = dict(
        means=f32(pts), scales=f32(scales), rotations=normal2rotation(f32(nrm)),
        opacities=f32(rng.normal(1.0, 1.0, N)), harmonics=f32(rng.random((N, 3))).reshape(N, 1, 3),
        view_scores=f32(rng.uniform(0, 2, N)), view_supports=f32(rng.poisson(3, N)),
        view_means=f32(vm))
    ext = torch.eye(4)[None]
    K = normalised_intrinsic(H, W, 60.0, 60.0)[None]
    return state, ext, K
