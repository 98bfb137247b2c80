"""Thin Python wrappers over the non-rasterizer entry points of libags_b200.so:
the fused image loss (K8) and the fused Adam (K7)."""
import ctypes as C
import torch

from . import lib as L


class LossResult:
    __slots__ = ("normal_unit", "d2n", "d_rgb", "d_normal", "d_depth", "terms", "workspace", "args", "keep", "gt_ptrs")

    def total(self, w_depth=0.8, w_cons=0.1, w_tv=0.1):
        t = self.terms
        return t[0] + w_depth * t[1] + w_cons * t[2] + w_tv * t[3]

    def frame_perf(self):
        """what track_performance stores: per-frame rgb-L1 mean + depth-L1 mean"""
        return self.terms[4::2] + self.terms[5::2]


def loss_forward_backward(rgb, normal, depth, opacity, rgb_gt, depth_gt, tanfov, *, B_total=None,
                          vis_count=None, w_depth=0.8, w_cons=0.1, w_tv=0.1, out=None, frame_weight=None,
                          want_maps=True):
    """Post-processing + loss + gradients for a batch of rendered frames (all (B,C,H,W) CUDA fp32).
    `tanfov` is (B,2) tan(fov/2).  `rgb_gt` / `depth_gt`: stacked (B,3,H,W) / (B,1,H,W) tensors, or LISTS
    of B per-frame tensors (3,H,W) / (1,H,W), read in place (no stacked copy).  `frame_weight` (B) float:
    0 marks a padded frame that contributes nothing.  `want_maps=False` skips the normal_unit / d2n outputs
    (the training loop never reads them).  Returns a LossResult; `terms` = [L_rgb, L_depth, L_cons, L_tv,
    (rgb_f, depth_f) per frame].  Passing `out` (a previous result) reuses all buffers and the cached
    argument struct (the training loop calls this with fixed pointers every iteration); with ground-truth
    lists the per-frame pointers are refreshed from the lists given."""
    lib = L.load()
    as_list = isinstance(rgb_gt, (list, tuple))
    if out is not None and getattr(out, "args", None) is not None:
        a = out.args
        if as_list:
            for k in range(a.B):
                out.gt_ptrs[0][k], out.gt_ptrs[1][k] = L.ptr(rgb_gt[k]), L.ptr(depth_gt[k])
            out.keep = out.keep[:7] + (list(rgb_gt), list(depth_gt))
        a.stream = L.current_stream(rgb.device)
        L.check(lib.ags_loss_forward_backward(C.byref(a)), "ags_loss_forward_backward")
        return out
    B, _, H, W = rgb.shape
    dev = rgb.device
    r = LossResult()
    o = dict(device=dev, dtype=torch.float32)
    r.normal_unit = torch.empty(B, 3, H, W, **o) if want_maps else None
    r.d2n = torch.empty(B, 3, H, W, **o) if want_maps else None
    r.d_rgb = torch.empty(B, 3, H, W, **o)
    r.d_normal = torch.empty(B, 3, H, W, **o)
    r.d_depth = torch.empty(B, 1, H, W, **o)
    r.terms = torch.empty(4 + 2 * B, **o)
    r.workspace = torch.empty(lib.ags_loss_scratch_bytes(B, H, W), device=dev, dtype=torch.uint8)
    a = L.LossArgs()
    a.B, a.H, a.W = B, H, W
    a.B_total = int(B_total or B)
    a.rgb, a.normal, a.depth, a.opacity = L.ptr(rgb), L.ptr(normal), L.ptr(depth), L.ptr(opacity)
    a.tanfov = L.ptr(tanfov)
    r.gt_ptrs = None
    if as_list:
        r.gt_ptrs = ((C.c_void_p * B)(), (C.c_void_p * B)())
        for k in range(B):
            r.gt_ptrs[0][k], r.gt_ptrs[1][k] = L.ptr(rgb_gt[k]), L.ptr(depth_gt[k])
        a.rgb_gt_frames_host = C.cast(r.gt_ptrs[0], C.POINTER(C.c_void_p))
        a.depth_gt_frames_host = C.cast(r.gt_ptrs[1], C.POINTER(C.c_void_p))
    else:
        a.rgb_gt, a.depth_gt = L.ptr(rgb_gt), L.ptr(depth_gt)
    a.vis_count = L.ptr(vis_count)
    a.frame_weight = L.ptr(frame_weight)
    a.normal_unit, a.d2n = L.ptr(r.normal_unit), L.ptr(r.d2n)
    a.d_rgb, a.d_normal, a.d_depth = L.ptr(r.d_rgb), L.ptr(r.d_normal), L.ptr(r.d_depth)
    a.loss_terms = L.ptr(r.terms)
    a.w_depth, a.w_cons, a.w_tv = w_depth, w_cons, w_tv
    a.workspace, a.workspace_bytes = L.ptr(r.workspace), r.workspace.numel()
    a.stream = L.current_stream(dev)
    r.args = a
    r.keep = (rgb, normal, depth, opacity, tanfov, vis_count, frame_weight,
              list(rgb_gt) if as_list else rgb_gt, list(depth_gt) if as_list else depth_gt)   # pointers stay valid
    L.check(lib.ags_loss_forward_backward(C.byref(a)), "ags_loss_forward_backward")
    return r


def adam_step(params, grads, exp_avgs, exp_avg_sqs, lrs, step=None, step_dev=None,
              betas=(0.9, 0.999), eps=1e-15, skip_flag_ptr=None, cache=None, zero_grad=False):
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
    a.zero_grad = int(bool(zero_grad))
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


# ------------------------------------------------------------------ per-keyframe map maintenance
MAP_FIELDS = (("means", 3), ("scales", 3), ("rotations", 4), ("opacities", 1), ("harmonics", 3),
              ("view_scores", 1), ("view_supports", 1), ("view_means", 3))


def _u8_scratch(pool, name, nbytes, device):
    """256-byte aligned scratch pointer of >= nbytes (from a buffer pool when given)."""
    t = pool.get(name, nbytes + 256) if pool is not None else torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    return t, (t.data_ptr() + 255) & ~255


def spawn(rgb, depth, depth_smooth, c2w, Kinv, pred, store, n_old, capacity, *, error_thres, voxel_size=0.02,
          seed=0, select_out=None, pool=None):
    """ags_spawn: append the new Gaussians of one RGB-D keyframe to the capacity buffers `store`
    (dict name -> (capacity, width) float32 CUDA tensors, MAP_FIELDS).  `pred` = (rgb, depth, opacity)
    render of the current map at the keyframe's pose or None.  c2w (4,4) / Kinv (3,3) are HOST
    tensors.  Returns (appended, candidates, wanted) -- one 16-byte D2H read, the only sync."""
    lib = L.load()
    dev = rgb.device
    _, H, W = rgb.shape
    a = L.SpawnArgs()
    a.H, a.W = H, W
    a.rgb, a.depth, a.depth_smooth = L.ptr(rgb), L.ptr(depth), L.ptr(depth_smooth)
    a.c2w = (C.c_float * 16)(*[float(x) for x in c2w.reshape(-1).tolist()])
    a.Kinv = (C.c_float * 9)(*[float(x) for x in Kinv.reshape(-1).tolist()])
    if pred is not None:
        a.pred_rgb, a.pred_depth, a.pred_opacity = L.ptr(pred[0]), L.ptr(pred[1]), L.ptr(pred[2])
    a.error_thres, a.voxel_size, a.seed = float(error_thres), float(voxel_size), int(seed) & 0xffffffff
    a.n_old, a.capacity = int(n_old), int(capacity)
    for name, _ in MAP_FIELDS:
        setattr(a, name, L.ptr(store[name]))
    counters = torch.empty(4, dtype=torch.int32, device=dev)
    a.counters = L.ptr(counters)
    a.select_out = L.ptr(select_out)
    keep, a.workspace = _u8_scratch(pool, "spawn", lib.ags_spawn_scratch_bytes(H, W), dev)
    a.workspace_bytes = lib.ags_spawn_scratch_bytes(H, W)
    a.stream = L.current_stream(dev)
    L.check(lib.ags_spawn(C.byref(a)), "ags_spawn")
    c = counters.tolist()
    return c[0], c[1], c[2]


def view_stats_update(count_last, means, rotations_raw, cam_pos, depth_max, use_view_distribution,
                      view_supports, view_means, view_scores):
    """ags_view_stats_update (mapping/gaussian_map.py:195-227), in place."""
    lib = L.load()
    N = means.shape[0]
    L.check(lib.ags_view_stats_update(N, L.ptr(count_last), L.ptr(means), L.ptr(rotations_raw), float(cam_pos[0]),
                                      float(cam_pos[1]), float(cam_pos[2]), float(depth_max),
                                      int(bool(use_view_distribution)), L.ptr(view_supports), L.ptr(view_means),
                                      L.ptr(view_scores), L.current_stream(means.device)), "ags_view_stats_update")


def prune_compact(src, dst, N, *, counts=None, prune_mask=None, pool=None):
    """ags_prune_compact: compact the eight SoA tensors of `src` (dict, MAP_FIELDS) into `dst`, dropping
    Gaussians flagged in `prune_mask` (uint8/bool (N), updated in place: quirk Q5), never counted in
    `counts` ((T,N) int32) or with sigmoid(opacity) < 0.1.  Returns the number kept (one 4-byte D2H)."""
    lib = L.load()
    dev = src["means"].device
    a = L.PruneArgs()
    a.N = int(N)
    a.T = int(counts.shape[0]) if counts is not None else 0
    a.counts = L.ptr(counts)
    if prune_mask is not None:
        assert prune_mask.dtype in (torch.uint8, torch.bool) and prune_mask.numel() == N
        a.prune_mask = L.ptr(prune_mask)
    for k, (name, _) in enumerate(MAP_FIELDS):
        a.src[k], a.dst[k] = L.ptr(src[name]), L.ptr(dst[name])
    n_kept = torch.empty(1, dtype=torch.int32, device=dev)
    a.n_kept = L.ptr(n_kept)
    need = lib.ags_prune_scratch_bytes(int(N))
    keep, a.workspace = _u8_scratch(pool, "prune", need, dev)
    a.workspace_bytes = need
    a.stream = L.current_stream(dev)
    L.check(lib.ags_prune_compact(C.byref(a)), "ags_prune_compact")
    return int(n_kept.item())


def view_utility(depth, confidence, voxel_centers, unexplored, w2c, K, depth_range, valid=None):
    """ags_view_utility: (explore (V,), exploit (V,)) of V rendered candidate views (planning/
    confidence.py:69-101, planning/exploration.py:62-86).  depth/confidence (V,h,w) float32; voxel_centers
    (M,3); unexplored (M) bool; w2c (V,4,4) inverse extrinsics; K (V,3,3) normalised intrinsics."""
    lib = L.load()
    dev = depth.device
    V, h, w = depth.shape
    a = L.UtilityArgs()
    a.V, a.h, a.w, a.M = V, h, w, int(voxel_centers.shape[0])
    un = unexplored.to(torch.uint8).contiguous() if unexplored.dtype != torch.uint8 else unexplored.contiguous()
    va = None if valid is None else valid.to(torch.uint8).contiguous()
    vc, wm, km = voxel_centers.float().contiguous(), w2c.float().contiguous(), K.float().contiguous()
    dc, cc = depth.float().contiguous(), confidence.float().contiguous()      # keep the temporaries alive
    a.depth, a.confidence, a.valid = L.ptr(dc), L.ptr(cc), L.ptr(va)
    a.voxel_centers, a.unexplored, a.w2c, a.K = L.ptr(vc), L.ptr(un), L.ptr(wm), L.ptr(km)
    a.depth_lo, a.depth_hi = float(depth_range[0]), float(depth_range[1])
    explore = torch.empty(V, dtype=torch.float32, device=dev)
    exploit = torch.empty(V, dtype=torch.float32, device=dev)
    a.explore, a.exploit = L.ptr(explore), L.ptr(exploit)
    a.stream = L.current_stream(dev)
    L.check(lib.ags_view_utility(C.byref(a)), "ags_view_utility")
    return explore, exploit


def voxel_roi(means, rotations_raw, opacities_raw, confidences, bbox_min, voxel_size, dim, *,
              confidence_thres=0.3, opacity_thres=0.7, min_gaussian_per_voxel=5):
    """ags_voxel_roi (mapping/voxel_map.py:70-113): per voxel of the (dim) grid the number of opaque
    low-confidence Gaussians, the normalised mean normal where that number exceeds
    min_gaussian_per_voxel (else 0) and the corresponding bool mask.  Returns (count (M,) int32,
    normal (M,3), mask (M,) bool)."""
    lib = L.load()
    dev = means.device
    N = means.shape[0]
    dim = [int(d) for d in dim]
    M = dim[0] * dim[1] * dim[2]
    a = L.VoxelRoiArgs()
    a.N = N
    keep = [means.float().contiguous(), rotations_raw.float().contiguous(), opacities_raw.float().contiguous(),
            confidences.float().contiguous()]
    a.means, a.rotations, a.opacities, a.confidences = [L.ptr(t) for t in keep]
    a.bbox_min = (C.c_float * 3)(*[float(x) for x in bbox_min])
    a.voxel_size = (C.c_float * 3)(*[float(x) for x in voxel_size])
    a.dim = (C.c_int32 * 3)(*dim)
    a.confidence_thres, a.opacity_thres = float(confidence_thres), float(opacity_thres)
    a.min_gaussian_per_voxel = int(min_gaussian_per_voxel)
    count = torch.empty(M, dtype=torch.int32, device=dev)
    normal = torch.empty(M, 3, dtype=torch.float32, device=dev)
    mask = torch.empty(M, dtype=torch.uint8, device=dev)
    a.voxel_count, a.voxel_normal, a.update_mask = L.ptr(count), L.ptr(normal), L.ptr(mask)
    a.stream = L.current_stream(dev)
    L.check(lib.ags_voxel_roi(C.byref(a)), "ags_voxel_roi")
    return count, normal, mask.bool()
