"""Host-side mirror of the reference's mapping/gaussian_map.py (class GaussianMap) on top of
libags_b200.so: same public surface -- update/train/save/load/get_attr/get_params, the get_*
properties, background_color/scene_near/scene_far/view_*/training_data/is_init -- so
mapping/mapper.py and the planners can use it unchanged (reference lines cited per method).

The per-iteration loop of train() (mapping/gaussian_map.py:76-127) is ONE fused device pipeline:
    render_forward (RAW params: activations fused, B views per launch)
      -> loss_forward_backward (post-processing + 4 loss terms + gradients, 2 kernels)
      -> render_backward (composite_bwd + project_bwd incl. activation backward)
      -> adam_step (5 groups, 1 kernel)
with every buffer allocated once per train() call and no autograd graph.  The only host<->device
traffic per iteration is the (B,) keyframe ids up and (B + stats) floats down for the
loss-weighted sampler, exactly the dependency the reference has (mapping/utils.py:206-218).
"""
import ctypes as C
import os
import numpy as np
import torch
import torch.nn.functional as F

from . import lib as L
from . import ops
from . import operations as O
from .rasterizer import RenderBatch, views_per_chunk


class WeightedSampler:
    """mapping/utils.py:190-228 (ids only; the frame tensors are gathered on the device)."""

    def __init__(self, cfg, n_frames):
        active = min(cfg.active_size, n_frames)
        ids = np.arange(n_frames)
        self.active_ids = ids[-active:]
        self.random_ids_all = ids[:-active]
        self.selected_num = min(len(self.random_ids_all), cfg.batch_size - active)
        self.v = len(self.active_ids) + self.selected_num

    def next_ids(self, weight_host):
        """`weight_host`: per-keyframe performance, numpy or torch (host).  When the batch takes EVERY
        non-active keyframe the draw is the whole population whatever the weights are; the 150 us
        np.random.choice over it is skipped then (the ids come back in index order, and numpy's global
        stream is not advanced -- the only place where the stream differs from the reference's)."""
        sel = self.active_ids.copy()
        if self.selected_num > 0:
            if self.selected_num == len(self.random_ids_all):
                return np.append(sel, self.random_ids_all)
            w = np.asarray(weight_host, dtype=np.float32)[self.random_ids_all]
            w = w / np.sum(w, dtype=np.float32)
            drawn = np.random.choice(self.random_ids_all, size=self.selected_num, p=w, replace=False)
            sel = np.append(sel, self.random_ids_all[drawn])
        return sel


class _BufferPool:
    """Device buffers that survive across train() calls and grow geometrically: the map gains
    Gaussians at every keyframe, so exact-size allocations would miss the caching allocator and pay
    a cudaMalloc (milliseconds) per buffer per keyframe."""

    def __init__(self, device):
        self.device, self.bufs = device, {}

    def get(self, name, nbytes):
        b = self.bufs.get(name)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes * 1.5) + 4096, dtype=torch.uint8, device=self.device)
            self.bufs[name] = b
        return b

    def floats(self, name, numel, zero=False):
        t = self.get(name, numel * 4)[:numel * 4].view(torch.float32)
        return t.zero_() if zero else t


_ATTR = {"means": "_means", "scales": "_scales", "rotations": "_rotations", "opacities": "_opacities",
         "harmonics": "_harmonics", "view_scores": "view_scores", "view_supports": "view_supports",
         "view_means": "view_means"}


class _MapStore:
    """Capacity buffers behind the eight SoA tensors of the map.  The reference re-allocates all of
    them with torch.cat at every keyframe and with boolean indexing at every prune
    (mapping/gaussian_map.py:403-462, 240-246); here the spawn kernel appends in place, the prune
    kernel compacts into a second set of buffers (ping-pong), and GaussianMap's attributes are
    [:N] views.  Tensors assigned from outside (load(), tests, the fused multi-GPU engine) are
    detected by their data pointers and copied back in."""

    def __init__(self, device):
        self.device, self.cap, self.buf, self.alt = device, 0, None, None

    def _alloc(self, cap):
        o = dict(device=self.device, dtype=torch.float32)
        return {n: (torch.empty(cap, w, **o) if w > 1 else torch.empty(cap, **o)) for n, w in ops.MAP_FIELDS}

    def owns(self, gm):
        if self.buf is None:
            return False
        return all(getattr(gm, _ATTR[n]).data_ptr() == self.buf[n].data_ptr() and
                   getattr(gm, _ATTR[n]).dtype == torch.float32 for n, _ in ops.MAP_FIELDS)

    def adopt(self, gm, extra):
        """make sure the map's tensors are prefix views of buffers with room for `extra` more rows"""
        N = gm._means.shape[0]
        if self.owns(gm) and N + extra <= self.cap:
            return
        if self.buf is not None and N + extra <= self.cap:
            # tensors assigned from outside (the symmetric flat buffer of the fused multi-GPU engine after
            # every train() call): copy home into the buffers we already have -- no allocation
            dst = self.buf
            if any(getattr(gm, _ATTR[n]).data_ptr() == dst[n].data_ptr() for n, _ in ops.MAP_FIELDS):
                dst = self.other()                     # partly aliased: go through the ping-pong set
        else:
            # grow geometrically (2x + room for the keyframe about to be added): a re-allocation is a
            # cudaMalloc per field, which with peer mappings enabled costs milliseconds per GPU of the box
            self.cap = max(2 * (N + extra) + 1024, 262144)
            # both halves of the ping-pong pair now: the first prune would otherwise pay eight cudaMallocs
            dst, self.alt = self._alloc(self.cap), self._alloc(self.cap)
        for n, w in ops.MAP_FIELDS:
            src = getattr(gm, _ATTR[n]).detach().reshape((N, w) if w > 1 else (N,))
            if src.data_ptr() != dst[n].data_ptr():
                dst[n][:N].copy_(src)
        if dst is self.alt:
            self.swap()
        else:
            self.buf = dst
        self.expose(gm, N)

    def expose(self, gm, n):
        for name, _ in ops.MAP_FIELDS:
            t = self.buf[name][:n]
            setattr(gm, _ATTR[name], t.view(n, 1, 3) if name == "harmonics" else t)

    def other(self):
        if self.alt is None:
            self.alt = self._alloc(self.cap)
        return self.alt

    def swap(self):
        self.buf, self.alt = self.alt, self.buf


class _TrainEngine:
    """Preallocated buffers + the fused iteration.  Everything that depends only on (B, H, W), the
    process group and where the keyframes live survives across train() calls (GaussianMap._engine: the
    reference calls train() once per keyframe with 10 iterations, so per-call set-up is a fixed cost of
    every update); bind() attaches the current map, whose N changes at every keyframe."""

    def __init__(self, gm, B, H, W, dist_ctx=None, on_host=False):
        dev = gm.device
        self.key = (B, H, W, id(dist_ctx), bool(on_host))
        self.gm, self.B, self.H, self.W, self.dev = gm, B, H, W, dev
        self.dist = dist_ctx
        self.world = dist_ctx.world if dist_ctx else 1
        self.fused = dist_ctx is not None and dist_ctx.fused
        self.on_host = bool(on_host)
        o = dict(device=dev, dtype=torch.float32)
        self.images = (torch.empty(B, 3, H, W, **o), torch.empty(B, 3, H, W, **o), torch.empty(B, 1, H, W, **o),
                       torch.empty(B, 1, H, W, **o), torch.empty(B, 1, H, W, **o))
        # host-resident keyframes (bench end-to-end mode): ground-truth staging is double buffered, the
        # H2D of step i+1 runs on a copy stream while backward/Adam of step i and the forward of step
        # i+1 execute (only the loss kernel reads the ground truth).  Device-resident keyframes are read
        # in place through per-frame pointers (no stacked copy).
        self.gt = None
        self.gt_k = 0
        if self.on_host:
            self.gt = [(torch.empty(B, 3, H, W, **o), torch.empty(B, 1, H, W, **o)) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream(device=dev)
            for a, b in self.gt:                       # written on the copy stream, freed on the compute stream
                a.record_stream(self.copy_stream); b.record_stream(self.copy_stream)
            self.copy_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.copy_pending = False
        self.prefetched = [{}, {}]          # per ground-truth buffer: local slot -> keyframe id already staged
        self.gt_lists = None                # device mode: ([rgb tensors], [depth tensors]) of the staged batch
        # camera blocks of the batch: one flat device buffer [B*16 view | B*16 proj | B*2 tanfov], gathered
        # by a kernel from the device table of all keyframes (ids travel as kernel arguments)
        self.cam_flat = torch.empty(B * 34, **o)
        self.view = self.cam_flat[:B * 16].view(B, 16)
        self.proj = self.cam_flat[B * 16:B * 32].view(B, 16)
        self.tanfov = self.cam_flat[B * 32:].view(B, 2)
        self.cam_ids = (C.c_int32 * B)()
        self.frame_w = torch.ones(B, **o)                # 0 marks a padded slot of a sharded batch
        self.frame_w_host = [1.0] * B
        self.vis_count = torch.empty(H, W, device=dev, dtype=torch.int32) if dist_ctx else None
        self.nterm = 2 * B + 4 + B + 2                # loss terms, per-frame perf | per-view instances | (instances, overflow)
        self.host = torch.empty(self.world * self.nterm, dtype=torch.float32).pin_memory()
        self.terms_all = torch.empty(self.world * self.nterm, **o) if dist_ctx else None
        self.terms_loc = torch.empty(self.nterm, **o) if dist_ctx else None
        self.host_stats = torch.empty(L.AGS_NUM_STATS, dtype=torch.int32).pin_memory()
        self.host_np, self.host_stats_np = self.host.numpy(), self.host_stats.numpy()
        self.event = torch.cuda.Event()
        self.loss_done = torch.cuda.Event()
        self.side_stream = torch.cuda.Stream(device=dev) if (dist_ctx is not None and os.environ.get("AGS_DIST_SIDE", "1") != "0") else None
        if dist_ctx is not None:
            dist_ctx.warm_up(dev)                    # NCCL's lazy channel set-up, outside any keyframe update
        self.aux = dist_ctx.aux_buffers(H * W, self.nterm, dev) if self.fused else None
        self.sync = dist_ctx.sync_buffers(dev) if (self.fused and dist_ctx.folded) else None
        self.sync_wait = None
        self.marks = [] if os.environ.get("AGS_DIST_PROFILE") else None     # (name, event) per segment boundary
        self.loss_outs = [None, None]       # one per ground-truth buffer (each has its own argument struct)
        self.loss_out = None
        self.vis_args = self.terms_args = None
        self.N = -1

    def bind(self, cam_table, B_total):
        """attach the map as it is now: parameter / gradient / Adam-state views, fresh Adam state
        (mapping/gaussian_map.py:259-292: a new optimiser per train() call), the rasterizer workspace"""
        gm, dev, B, H, W = self.gm, self.dev, self.B, self.H, self.W
        N = gm._means.shape[0]
        self.N, self.B_total = N, int(B_total)
        self.cam_table = cam_table                       # (T, 34) device: view | proj | tanfov per keyframe
        o = dict(device=dev, dtype=torch.float32)
        self.params = [gm._means, gm._scales, gm._rotations, gm._opacities, gm._harmonics]
        for p in self.params:
            assert p.is_contiguous() and p.dtype == torch.float32
        total = sum(p.numel() for p in self.params)
        self.flat = None
        self.m = self.v = None
        if self.fused:
            # parameters and gradients live in symmetric (peer-mapped) flat buffers; the GaussianMap
            # tensors become views of the parameter buffer for the duration of train()
            self.flat = self.dist.flat_buffers(total, dev)
            off, views = 0, []
            for p in self.params:
                dst = self.flat.param[off:off + p.numel()].view(p.shape)
                if dst.data_ptr() != p.data_ptr():
                    dst.copy_(p)
                views.append(dst)
                off += p.numel()
            gm._means, gm._scales, gm._rotations, gm._opacities, gm._harmonics = views
            self.params = views
            self.grad_flat = self.flat.grad
            # the exchange covers the live 14*N floats rounded up to 4*world (one 16-byte element per rank),
            # not the capacity of the symmetric buffers
            q = 4 * self.world
            self.active_padded = (total + q - 1) // q * q
            self.m_flat = gm._pool.floats("adam_m", self.active_padded, zero=True)
            self.v_flat = gm._pool.floats("adam_v", self.active_padded, zero=True)
        else:
            # gradients are views of ONE flat buffer (14 floats per Gaussian): a single all-reduce in the
            # NCCL-baseline sharded path.  Zeroed once here: the Adam kernel re-zeroes what it consumed.
            self.grad_flat = gm._pool.floats("grad", total, zero=True)
            mf, vf = gm._pool.floats("adam_m", total, zero=True), gm._pool.floats("adam_v", total, zero=True)
            self.m, self.v, off = [], [], 0
            for p in self.params:
                self.m.append(mf[off:off + p.numel()].view(p.shape))
                self.v.append(vf[off:off + p.numel()].view(p.shape))
                off += p.numel()
        self.grads, off = [], 0
        for p in self.params:
            self.grads.append(self.grad_flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.lrs = [gm.cfg.optimizer.mean_lr, gm.cfg.optimizer.scale_lr, gm.cfg.optimizer.rotation_lr,
                    gm.cfg.optimizer.opacity_lr, gm.cfg.optimizer.harmonic_lr]
        self.step = 0
        self.conf = gm.get_confidences.contiguous()
        self.bg = gm.background_color.to(dev).float().contiguous()
        cap = gm._inst_cap_hint(N, B)
        self.rb = RenderBatch(gm._means, gm._scales, gm._rotations, gm._opacities,
                              gm._harmonics.reshape(N, 3), self.conf, self.view, self.proj, self.tanfov,
                              self.bg, H, W, param_mode=L.PARAMS_RAW, scale_factor=gm.scale_factor,
                              scale_max=0.05, inst_cap=cap, with_importance=False,
                              pool=lambda nbytes: gm._pool.get("workspace", nbytes), images=self.images)
        # RenderBatch copies nothing for contiguous fp32 inputs, but make the aliasing explicit
        self.rb.inputs = [gm._means, gm._scales, gm._rotations, gm._opacities,
                          gm._harmonics.reshape(N, 3), self.conf]
        self.rb.view = [self.view, self.proj, self.tanfov, self.bg]
        if self.fused:
            self.rb.stats = self.flat.stats          # peers read the overflow flag through NVLink
        self.fwd_args = self.rb._args()      # argument structs are built once per bind: pointers never change
        self.grad_args = [None, None]
        self.adam_cache = {}
        self.dist_args = None
        for lo in self.loss_outs:
            if lo is not None:
                lo.args.B_total = self.B_total
        self.copy_pending = False
        self.prefetched = [{}, {}]

    def set_batch(self, frames, idx, weights=None, ids=None):
        """Stage this iteration's keyframes (`frames`: list of B (rgb (3,H,W), depth (1,H,W)) pairs) and
        camera blocks (rows `idx` of the device camera table).  Device-resident keyframes are only
        referenced (the loss kernel reads them in place); pinned host frames are uploaded into the
        double-buffered staging tensors -- the per-step H2D of the host-resident mode.  `weights` (B):
        0 for padded slots.  `ids` (keyframe ids of the batch) lets the copy skip frames that
        prefetch_next() already staged in this buffer."""
        B = self.B
        # camera blocks: gathered on the device from the table of all keyframes; the ids travel as
        # kernel arguments (an H2D copy here would queue on the copy engine behind the keyframe
        # uploads and stall the forward)
        for k in range(B):
            self.cam_ids[k] = int(idx[k])
        L.check(L.load().ags_stage_cameras(self.cam_table.data_ptr(), self.cam_table.shape[0], self.cam_ids, B,
                                           self.view.data_ptr(), self.proj.data_ptr(), self.tanfov.data_ptr(),
                                           L.current_stream(self.dev)), "ags_stage_cameras")
        wl = [1.0] * B if weights is None else [float(x) for x in weights]
        if wl != self.frame_w_host:
            self.frame_w.copy_(torch.tensor(wl, dtype=torch.float32), non_blocking=True)
            self.frame_w_host = wl
        if not self.on_host:
            self.gt_lists = ([f[0] for f in frames], [f[1] for f in frames])
            return
        self.gt_k ^= 1
        rgb_gt, depth_gt = self.gt[self.gt_k]
        # buffer gt_k was last read by the loss of step i-2, which finished before fetch(i-1)
        staged = self.prefetched[self.gt_k]
        with torch.cuda.stream(self.copy_stream):
            for k in range(B):
                if ids is not None and staged.get(k) == int(ids[k]):
                    continue
                rgb_gt[k].copy_(frames[k][0], non_blocking=True)
                depth_gt[k].copy_(frames[k][1], non_blocking=True)
            self.copy_done[self.gt_k].record(self.copy_stream)
        staged.clear()
        self.copy_pending = True

    def prefetch_next(self, fixed, frame_of):
        """Host-resident mode: start the H2D of the NEXT step's keyframes whose ids do not depend on the
        sampler draw (`fixed`: local batch slot -> keyframe id; the sampler's always-selected active
        frames, mapping/utils.py:196-204) into the other ground-truth buffer.  Called right after a step
        is enqueued and before the host waits for its loss terms, so this part of the per-step copy
        overlaps the forward + loss of the current step; the drawn frames follow in set_batch().  The
        target buffer was last read by the loss of the previous step, which the host has already waited
        for."""
        if not fixed or not self.on_host:
            return
        nk = self.gt_k ^ 1
        staged = self.prefetched[nk]
        if staged:
            return                                   # already staged (overflow retry of the same step)
        rgb_gt, depth_gt = self.gt[nk]
        with torch.cuda.stream(self.copy_stream):
            for k, fid in fixed.items():
                rgb, depth = frame_of(fid)
                rgb_gt[k].copy_(rgb, non_blocking=True)
                depth_gt[k].copy_(depth, non_blocking=True)
                staged[k] = int(fid)

    def drain(self):
        """nothing of this engine may still be in flight on the copy stream when train() returns"""
        if self.on_host:
            self.copy_stream.synchronize()
            self.prefetched = [{}, {}]

    def grow(self, need):
        """re-plan the instance capacity after an overflow (nothing was rendered or updated)"""
        self.rb.inst_cap = min(int(need * 1.3) + 65536, 2 ** 31 - 1)
        self.rb._alloc(L.load())
        self.fwd_args = self.rb._args()

    def iterate(self):
        """Enqueue one optimisation step on the staged batch.  Order on the stream:
        forward -> loss(+grads) -> [tiny D2H of loss terms / perf / stats + event] -> backward ->
        Adam.  The host later waits on the event only, so the backward and the Adam step overlap
        with the host preparing the next batch.  Adam is gated on the device overflow flag."""
        lib = L.load()
        rb = self.rb
        st = L.current_stream(self.dev)
        fa = self.fwd_args
        fa.stream = st
        if self.sync is not None:
            self.sync.epoch += 1                      # one epoch per enqueued iteration (retries included)
        self._mark("start")
        L.check(lib.ags_render_forward(C.byref(fa)), "ags_render_forward")
        self._mark("forward")
        vis = None
        if self.fused:
            vis = self._fused_vis_count(lib, st)
        elif self.dist is not None:
            # every rank must take the same overflow decision: (instances, overflow) -> MAX.
            # (the fused path reads the peers' flags itself; the host gets them via the gather)
            self.dist.all_reduce_max_(rb.stats[:2])
            seen = (rb.opacity[:, 0] > 1e-3) & (self.frame_w[:, None, None] > 0)
            torch.sum(seen, dim=0, dtype=torch.int32, out=self.vis_count)
            self.dist.all_reduce_sum_(self.vis_count)
            vis = self.vis_count
        k = self.gt_k if self.on_host else 0
        if self.on_host:
            if self.copy_pending:
                torch.cuda.current_stream(self.dev).wait_event(self.copy_done[k])
            gts = self.gt[k]
        else:
            gts = self.gt_lists
        self.loss_outs[k] = ops.loss_forward_backward(
            rb.rgb, rb.normal, rb.depth, rb.opacity, gts[0], gts[1], self.tanfov,
            B_total=self.B_total, vis_count=vis, out=self.loss_outs[k], frame_weight=self.frame_w, want_maps=False)
        lo = self.loss_out = self.loss_outs[k]
        self._mark("loss")
        if self.fused and self.sync is not None and self.side_stream is not None:
            # the exchange of the loss terms (a one-block kernel, a flag wait and the small D2H the host waits for) does
            # not feed the backward: it runs on a side stream next to it instead of between the loss and the backward
            main = torch.cuda.current_stream(self.dev)
            self.loss_done.record(main)
            with torch.cuda.stream(self.side_stream):
                self.side_stream.wait_event(self.loss_done)
                self._fused_terms_gather(lib, L.current_stream(self.dev), lo)
                self.host_stats.copy_(rb.stats, non_blocking=True)
                self.event.record(self.side_stream)
        elif self.fused:
            self._fused_terms_gather(lib, st, lo)
        elif self.dist is not None:
            # loss terms + per-frame performance of every rank, gathered on the stream before the
            # backward is enqueued: the host waits for this small copy only
            nt = 4 + 2 * self.B
            self.terms_loc[:nt].copy_(lo.terms)
            self.terms_loc[nt:nt + self.B].copy_(rb.stats[L.STAT_VIEW0:L.STAT_VIEW0 + self.B])
            self.terms_loc[self.nterm - 2:].copy_(rb.stats[:2])
            self.dist.all_gather_into_(self.terms_all, self.terms_loc)
            self.host.copy_(self.terms_all, non_blocking=True)
        else:
            self.host[:4 + 2 * self.B].copy_(lo.terms, non_blocking=True)
        if not (self.fused and self.sync is not None and self.side_stream is not None):
            self.host_stats.copy_(rb.stats, non_blocking=True)
            self.event.record(torch.cuda.current_stream(self.dev))
        single = self.dist is None
        if self.grad_args[k] is None:
            g = L.RenderGradArgs()
            g.d_rgb, g.d_normal, g.d_depth = L.ptr(lo.d_rgb), L.ptr(lo.d_normal), L.ptr(lo.d_depth)
            g.d_opacity = g.d_confidence = None
            (g.d_means3D, g.d_scales, g.d_rotations, g.d_opacities, g.d_colors) = [
                L.ptr(t) for t in self.grads]
            g.d_means2D = None
            # single GPU: the gradients start at zero (bind) and the Adam kernel zeroes them again as it
            # consumes them, so the backward accumulates and no separate zeroing pass runs
            g.accumulate = 1 if single else 0
            g.clear_records = 0                 # exactly one backward per forward in this loop
            self.grad_args[k] = g
        L.check(lib.ags_render_backward(C.byref(fa), C.byref(self.grad_args[k])), "ags_render_backward")
        self._mark("backward")
        self.step += 1
        if self.fused:
            self._fused_exchange_and_adam(lib, st)
            return
        if self.dist is not None:
            self.dist.all_reduce_grads_(self.grads)
        ops.adam_step(self.params, self.grads, self.m, self.v, self.lrs, step=self.step,
                      skip_flag_ptr=rb.stats.data_ptr() + 4 * L.STAT_OVERFLOW, cache=self.adam_cache,
                      zero_grad=single)

    def _mark(self, name):
        if self.marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.dev))
            self.marks.append((name, e))

    def segment_times(self):
        """AGS_DIST_PROFILE=1: median device time (us) between consecutive marks, by segment name"""
        torch.cuda.synchronize()
        acc = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            acc.setdefault("host gap" if n1 == "start" else n1, []).append(e0.elapsed_time(e1) * 1e3)
        return {k: round(float(np.median(v)), 1) for k, v in acc.items()}

    def _fused_vis_count(self, lib, st):
        """quirk Q1 over the whole batch: local count -> cross-GPU barrier -> sum of all ranks' planes
        over NVLink (csrc/dist_loss.cu), no NCCL collective"""
        a, d, x = self.vis_args, self.dist, self.aux
        if a is None:
            a = L.DistVisArgs()
            a.world, a.rank, a.B, a.H, a.W = d.world, d.rank, self.B, self.H, self.W
            a.opacity, a.vis_local, a.vis_count = L.ptr(self.images[3]), L.ptr(x.vis), L.ptr(self.vis_count)
            a.frame_weight = L.ptr(self.frame_w)
            for p in range(d.world):
                a.vis_peers[p] = x.vis_ptrs[p]
            a.vis_multicast = x.vis_mc if (d.use_multicast and x.vis_mc) else None
            self._fill_sync(a.sync)
            self.vis_args = a
        a.stream = st
        if self.sync is not None:
            a.sync.epoch = self.sync.epoch
        L.check(lib.ags_dist_vis_local(C.byref(a)), "ags_dist_vis_local")
        self._mark("vis local")
        if self.sync is None:
            x.barrier()
            self._mark("barrier A")
        L.check(lib.ags_dist_vis_sum(C.byref(a)), "ags_dist_vis_sum")
        self._mark("vis sum")
        return self.vis_count

    def _fused_terms_gather(self, lib, st, lo):
        """every rank's loss terms / per-frame performance / (instances, overflow) into every rank's
        gather buffer with peer stores, then the small D2H the host waits for"""
        a, d, x = self.terms_args, self.dist, self.aux
        if a is None:
            a = L.DistTermsArgs()
            a.world, a.rank, a.nterm, a.nview = d.world, d.rank, self.nterm, self.B
            self._fill_sync(a.sync)
            self.terms_args = a
        # the gather buffer is double buffered by iteration parity: a fast peer may already push the terms of the
        # next iteration while this rank's D2H of the current ones is still queued on the side stream
        half = (self.sync.epoch & 1) if self.sync is not None else 0
        off = half * d.world * self.nterm * 4
        for p in range(d.world):
            a.gather_peers[p] = x.gather_ptrs[p] + off
        a.gather_multicast = (x.gather_mc + off) if (d.use_multicast and x.gather_mc) else None
        a.stats = L.ptr(self.rb.stats)
        a.terms = L.ptr(lo.terms)
        a.stream = st
        if self.sync is not None:
            a.sync.epoch = self.sync.epoch
        L.check(lib.ags_dist_terms_put(C.byref(a)), "ags_dist_terms_put")
        if self.sync is None:
            x.barrier()
        else:
            self._wait(lib, L.SYNC_TERMS, st)          # every rank's terms have landed in the local gather buffer
        if self.sync is None:
            self._mark("terms put + barrier T")        # (folded mode: this runs on the side stream, off the critical path)
        n = d.world * self.nterm
        self.host.copy_(x.gather[half * n:(half + 1) * n], non_blocking=True)

    def _fused_exchange_and_adam(self, lib, st):
        """reduce-scatter(grads) -> Adam(shard) -> all-gather(params) in one kernel, between two
        device-side cross-GPU barriers on this stream (csrc/dist_adam.cu)"""
        f, d = self.flat, self.dist
        a = self.dist_args
        if a is None:
            a = L.DistAdamArgs()
            a.world, a.rank, a.num_groups = d.world, d.rank, len(self.params)
            for p in range(d.world):
                a.grad_peers[p], a.param_peers[p] = f.grad_ptrs[p], f.param_ptrs[p]
                a.skip_peers[p] = f.stats_ptrs[p] + 4 * L.STAT_OVERFLOW           # any rank's flag skips the step
            use_mc = d.use_multicast and f.grad_mc != 0 and f.param_mc != 0
            a.grad_multicast = f.grad_mc if use_mc else None
            a.param_multicast = f.param_mc if use_mc else None
            a.exp_avg, a.exp_avg_sq = L.ptr(self.m_flat), L.ptr(self.v_flat)
            for k, p in enumerate(self.params):
                a.numel[k] = p.numel()
                a.lr[k] = self.lrs[k]
            a.numel_padded = self.active_padded
            a.beta1, a.beta2, a.eps = 0.9, 0.999, 1e-15
            self._fill_sync(a.sync)
            self.dist_args = a
        a.step = self.step
        a.stream = st
        if self.sync is None:
            f.barrier()                                  # every rank's gradients are complete
            self._mark("barrier G1")
            L.check(lib.ags_dist_adam_step(C.byref(a)), "ags_dist_adam_step")
            self._mark("dist adam")
            f.barrier()                                  # every rank's parameters are updated
            self._mark("barrier G2")
            return
        # folded: the kernel signals "my gradients are complete" on entry, waits for all ranks, and signals
        # "my shard is in everybody's parameter buffer" from its last block; the next forward (or the end of
        # training) waits for that signal of all ranks
        a.sync.epoch = self.sync.epoch
        L.check(lib.ags_dist_adam_step(C.byref(a)), "ags_dist_adam_step")
        self._mark("dist adam")
        self._wait(lib, L.SYNC_PARAMS, st)
        self._mark("wait params")

    def _fill_sync(self, sy):
        if self.sync is None:
            return
        for p in range(self.dist.world):
            sy.peers[p] = self.sync.ptrs[p]
        sy.epoch = self.sync.epoch

    def _wait(self, lib, phase, st):
        sy = self.sync_wait
        if sy is None:
            sy = self.sync_wait = L.DistSync()
            self._fill_sync(sy)
        sy.epoch = self.sync.epoch
        L.check(lib.ags_dist_wait(C.byref(sy), phase, self.dist.world, self.dist.rank, st), "ags_dist_wait")

    def fetch(self):
        """Wait for the loss terms / per-frame performance / instance statistics of the step that
        was just enqueued (the sampler needs them, mapping/utils.py:206-218).  Plain numpy on the pinned
        result buffers: this sits on the host's critical path between two iterations."""
        B = self.B
        self.event.synchronize()
        h = self.host_np.reshape(-1, self.nterm)            # one row per rank (views of the pinned buffer)
        terms = h[:, :4].sum(0)                             # every rank's terms are already / B_total
        pf = h[:, 4:4 + 2 * B]
        perf = (pf[:, 0::2] + pf[:, 1::2]).reshape(-1)      # ordered like the (padded) batch; a fresh array
        nt = 4 + 2 * B
        if self.dist is not None:                           # global view: max instances, any overflow
            instances = int(h[:, self.nterm - 2].max())
            overflow = int(h[:, self.nterm - 1].max())
            view_cost = h[:, nt:nt + B].reshape(-1).copy()  # instances per frame, ordered like the batch
        else:
            st = self.host_stats_np
            instances, overflow = int(st[L.STAT_INSTANCES]), int(st[L.STAT_OVERFLOW])
            view_cost = st[L.STAT_VIEW0:L.STAT_VIEW0 + B].astype(np.float32)
        return terms, perf, (instances, overflow, int(self.host_stats_np[L.STAT_VISIBLE])), view_cost


class GaussianMap:
    """mapping/gaussian_map.py:17-590."""

    def __init__(self, cfg, device):
        self.device = torch.device(device)
        dev = self.device
        self._means = torch.empty(0, 3, device=dev)
        self._scales = torch.empty(0, 3, device=dev)
        self._rotations = torch.empty(0, 4, device=dev)
        self._opacities = torch.empty(0, device=dev)
        self._harmonics = torch.empty(0, 1, 3, device=dev)
        self.view_scores = torch.empty(0, device=dev)
        self.view_supports = torch.empty(0, device=dev)
        self.view_means = torch.empty((0, 3), device=dev)
        self.training_performance = torch.tensor([], device=dev)
        self.training_data = []
        self.is_init = False
        self.use_view_distribution = True
        self.frames_on_host = False          # bench end-to-end mode: keyframes stay in pinned host memory
        self.dist = None                     # active_gs_b200.distributed.FrameShard or None
        self.last_train_log = []
        self._cap_per_gaussian = 4.0
        self._pool = _BufferPool(self.device)
        self._store = _MapStore(self.device)
        self._cams = []                      # per keyframe: host camera products, computed once
        self._gt_cache = []                  # per keyframe: (source rgb, source depth, training rgb, training depth)
        self._engine = None                  # _TrainEngine, kept across train() calls
        self._frame_cost = {}                # keyframe id -> instances of its last training render
        if cfg is not None:
            self.cfg = cfg
            self.use_view_distribution = cfg.use_view_distribution
            self.scene_near, self.scene_far = cfg.bound
            self.sparse_ratio = cfg.sparse_ratio
            self.scale_factor = cfg.scale_factor
            self.error_thres = cfg.error_thres
            self.prune_interval = cfg.prune_interval
            self.optimization_steps = cfg.optimization_steps
            self.background_color = torch.tensor(cfg.background, dtype=torch.float32).to(dev)

    # ------------------------------------------------------------------ public loop (:62-130)
    def update(self, dataframe):
        self.add_gaussians(dataframe)
        self.train()

    def _inst_cap_hint(self, N, B):
        return int(self._cap_per_gaussian * N * B) + 65536

    def _camera(self, i):
        """Host-side camera products of keyframe i (utils/operations.py:748-762 for one view),
        computed once per keyframe instead of once per view per iteration: (34,) row
        [viewmatrix 16 | projmatrix 16 | tanfov 2], camera position, c2w, K^-1."""
        f = self.training_data[i]
        while len(self._cams) <= i:
            self._cams.append(None)
        c = self._cams[i]
        if c is None or c["ext"] is not f["extrinsic"] or c["K"] is not f["intrinsic"]:
            ext = f["extrinsic"].detach().float().cpu()
            K = f["intrinsic"].detach().float().cpu()
            _, view, proj, campos, tanfov = O.camera_blocks(ext[None], K[None], (self.scene_near, self.scene_far))
            c = dict(ext=f["extrinsic"], K=f["intrinsic"], c2w=ext, Kinv=torch.linalg.inv(K), campos=campos[0],
                     row=torch.cat([view.reshape(16), proj.reshape(16), tanfov.reshape(2)]))
            self._cams[i] = c
        return c

    def _camera_rows(self, ids):
        """(len(ids), 34) host tensor of camera rows"""
        return torch.stack([self._camera(i)["row"] for i in ids])

    def _frame_gt(self, i):
        """(rgb (3,H,W), depth (1,H,W)) of keyframe i as the training loop reads them: contiguous fp32, on
        the device unless the map keeps its keyframes in pinned host memory (frames_on_host).  Converted
        copies are cached per keyframe (the reference moves every dataframe to the device once,
        mapping/mapper.py:95)."""
        f = self.training_data[i]
        while len(self._gt_cache) <= i:
            self._gt_cache.append(None)
        c = self._gt_cache[i]
        if c is not None and c[0] is f["rgb"] and c[1] is f["depth"]:
            return c[2], c[3]

        def conv(t):
            if self.frames_on_host:
                t = t.detach().float().contiguous().cpu()
                return t if t.is_pinned() else t.pin_memory()
            return t.detach().to(self.device, torch.float32).contiguous()

        rgb, depth = conv(f["rgb"]), conv(f["depth"])
        self._gt_cache[i] = (f["rgb"], f["depth"], rgb, depth)
        return rgb, depth

    def begin_training(self):
        """Everything train() sets up once per call (mapping/gaussian_map.py:71-74): fresh Adam
        state, the sampler, the camera table of all keyframes, and the (persistent) engine."""
        from types import SimpleNamespace
        T = len(self.training_data)
        self._make_contiguous()
        sampler = WeightedSampler(self.cfg.sampler, T)
        B = sampler.v if self.dist is None else self.dist.local_batch(sampler.v)
        _, H, W = self.training_data[0]["rgb"].shape
        cam_rows = self._camera_rows(range(T))                       # host (T, 34), cached per keyframe
        # batch slots whose keyframe never changes (the sampler's active frames come first in the
        # sampled ids): local slot k of this rank is global slot rank*B + k
        if self.dist is None:
            fixed = {k: int(sampler.active_ids[k]) for k in range(min(B, len(sampler.active_ids)))}
        else:                                           # active keyframes are pinned round robin (FrameShard.balance)
            fixed = {}
            for j, fid in enumerate(sampler.active_ids[:sampler.v]):
                r, k = self.dist.pinned_slot(j)
                if r == self.dist.rank:
                    fixed[k] = int(fid)
        on_host = bool(self.frames_on_host)
        key = (B, H, W, id(self.dist), on_host)
        eng = self._engine
        if eng is None or eng.key != key:
            eng = self._engine = _TrainEngine(self, B, H, W, self.dist, on_host=on_host)
        eng.bind(cam_rows.to(self.device), sampler.v)
        perf_host = self.training_performance.detach().float().cpu().clone()
        return SimpleNamespace(fixed=fixed, sampler=sampler, B=B, H=H, W=W, eng=eng,
                               perf_host=perf_host, perf_np=perf_host.numpy(), log=[])

    def train_step(self, ctx, ids=None):
        """One iteration of mapping/gaussian_map.py:76-127: sample keyframes, stage them, enqueue
        forward/loss/backward/Adam, wait for the loss terms (sampler dependency)."""
        eng = ctx.eng
        sampled = ids is None
        ids = np.asarray(ids) if ids is not None else ctx.sampler.next_ids(ctx.perf_np)
        weights = None
        if self.dist is not None:
            # same keyframes on every rank; the batch is padded to a multiple of the world size (-1 =
            # padding: rendered from a repeated keyframe with loss weight 0) and, when the sampler drew
            # it, partitioned over the ranks by cost (instances of the last render)
            if sampled:
                ids = self.dist.balance(ids, len(ctx.sampler.active_ids), self._frame_cost)
            else:
                ids = self.dist.pad(ids)
            my = self.dist.my_frames(ids)
            weights = [1.0 if i >= 0 else 0.0 for i in my]
            filler = int(ids[ids >= 0][0])
            my = np.asarray([i if i >= 0 else filler for i in my])
        else:
            my = ids
        eng.set_batch([self._frame_gt(int(i)) for i in my], my, weights=weights, ids=my)
        while True:
            eng.iterate()
            if sampled:
                eng.prefetch_next(ctx.fixed, self._frame_gt)
            terms, perf, (instances, overflow, visible), view_cost = eng.fetch()
            if overflow == 0:
                break
            # capacity exceeded: nothing was rendered and the device-side flag turned the Adam
            # step into a no-op, so grow the workspace and redo the same iteration
            eng.step -= 1
            eng.grow(instances)
        need = float(instances) / max(1, eng.N * ctx.B)
        self._cap_per_gaussian = max(self._cap_per_gaussian, 1.5 * need)
        if self.dist is not None:
            real = ids >= 0
            rid, perf, view_cost = ids[real], perf[real], view_cost[real]
        else:
            rid = ids
        ctx.perf_np[rid] = perf
        self._frame_cost.update(zip(rid.tolist(), view_cost.tolist()))
        loss = float(terms[0] + 0.8 * terms[1] + 0.1 * terms[2] + 0.1 * terms[3])
        ctx.log.append((loss, perf, instances, visible))
        return loss

    def end_training(self, ctx):
        ctx.eng.drain()
        if ctx.eng.fused:
            # the parameters live in the symmetric flat buffer while training (sub-tensors start at arbitrary
            # 4-byte offsets there): bring them home into the capacity buffers the per-keyframe kernels expect
            self._store.adopt(self, 0)
        self.training_performance = ctx.perf_host.to(self.device)
        self.last_train_log = ctx.log

    def train(self, steps=None, frame_id_batches=None):
        """mapping/gaussian_map.py:66-130.  `frame_id_batches` (optional) overrides the sampler with
        explicit keyframe ids per iteration (parity tests)."""
        iterations = self.optimization_steps if steps is None else steps
        ctx = self.begin_training()
        for it in range(iterations):
            self.train_step(ctx, None if frame_id_batches is None else frame_id_batches[it])
        self.end_training(ctx)
        self.post_processing()
        self.is_init = True

    def _make_contiguous(self):
        for n in ["_means", "_scales", "_rotations", "_opacities", "_harmonics"]:
            setattr(self, n, getattr(self, n).detach().float().contiguous())

    # ------------------------------------------------------------------ :141-246
    POST_CHUNK = 4           # views per count-render launch when all T keyframes are re-rendered: the transient
                             # workspace is ~80 MB per view at 330 k surfels, and the first prune pass of a mission pays
                             # for allocating it (16 views: 1.4 GB, ~60 ms once; 4 views: four launches, ~0.2 ms more)

    def _render_raw(self, ids, H, W, *, render_mask=None, require_importance=False, front_only=False,
                    with_confidence=False):
        """Forward-only render of keyframes `ids` straight from the RAW parameters (activations fused
        in the projection kernel: no get_attr() pass), cameras from the per-keyframe cache."""
        rows = self._camera_rows(ids).to(self.device)
        B = len(ids)
        N = self._means.shape[0]
        rb = RenderBatch(self._means, self._scales, self._rotations, self._opacities,
                         self._harmonics.reshape(N, 3), self.get_confidences if with_confidence else None,
                         rows[:, :16], rows[:, 16:32], rows[:, 32:34], self.background_color, H, W,
                         render_mask=render_mask, require_importance=require_importance, front_only=front_only,
                         param_mode=L.PARAMS_RAW, scale_factor=self.scale_factor, scale_max=0.05,
                         pool=lambda nbytes: self._pool.get("workspace_aux", nbytes))
        rb.forward(check_overflow=True)
        return rb

    def post_processing(self):
        """:141-232.  One count-only render of the newest keyframe (all T keyframes every
        prune_interval-th time, in chunks) from the raw parameters, the confidence bookkeeping in one
        kernel (ags_view_stats_update) and the prune as one ordered compaction (ags_prune_compact).
        With a process group the T-1 older keyframes of a prune pass are sharded round robin over the ranks and
        the per-Gaussian "counted in some view" sums are all-reduced (4 B per Gaussian): the reference renders
        all T frames every 5th keyframe, T grows without bound."""
        T = len(self.training_data)
        require_prune = T % self.prune_interval == 0
        self._make_contiguous()
        _, H, W = self.training_data[-1]["depth"].shape
        N = self._means.shape[0]
        seen = None                                   # (1,N) int32: times counted over the keyframes before the newest
        if require_prune and T > 1:
            older = list(range(T - 1))
            if self.dist is not None:
                older = older[self.dist.rank::self.dist.world]
            step = views_per_chunk(N, H, W, per_gaussian=self._cap_per_gaussian, cap=self.POST_CHUNK)
            seen = torch.zeros(1, N, device=self.device, dtype=torch.int32)
            for c0 in range(0, len(older), step):
                chunk = older[c0:c0 + step]
                dgt = torch.stack([self.training_data[i]["depth"] for i in chunk]).to(self.device)
                rb = self._render_raw(chunk, H, W, render_mask=(dgt > 0.0).float(), require_importance=True,
                                      front_only=True)
                seen += rb.count.sum(0, keepdim=True, dtype=torch.int32)
            if self.dist is not None:
                self.dist.all_reduce_sum_(seen)
        dgt = self.training_data[T - 1]["depth"][None].to(self.device)
        rb = self._render_raw([T - 1], H, W, render_mask=(dgt > 0.0).float(), require_importance=True, front_only=True)
        last = rb.count                               # (1,N): the newest keyframe, rendered by every rank
        for n in ["view_scores", "view_supports", "view_means"]:
            setattr(self, n, getattr(self, n).float().contiguous())
        ops.view_stats_update(last[-1], self._means, self._rotations, self._camera(T - 1)["campos"],
                              float(self.training_data[-1]["depth_range"][1]), self.use_view_distribution,
                              self.view_supports, self.view_means, self.view_scores)
        if require_prune:
            counts = last if seen is None else torch.cat([seen, last], 0).contiguous()
            self._compact(counts=counts)

    def _compact(self, counts=None, prune_mask=None):
        N = self._means.shape[0]
        st = self._store
        st.adopt(self, 0)
        n = ops.prune_compact(st.buf, st.other(), N, counts=counts, prune_mask=prune_mask, pool=self._pool)
        st.swap()
        st.expose(self, n)
        print(f"delete {N - n} gaussians")

    def prune(self, prune_mask):
        """:234-246.  `prune_mask` (N,) bool is OR-ed in place with opacity < 0.1 (quirk Q5)."""
        if prune_mask.dtype not in (torch.bool, torch.uint8) or not prune_mask.is_contiguous():
            raise TypeError("prune_mask must be a contiguous bool tensor")
        self._make_contiguous()
        self._compact(prune_mask=prune_mask)

    # ------------------------------------------------------------------ :294-489
    def add_gaussians(self, dataframe):
        """:294-468 on the device: bilateral filter (ags_smooth_depth), render of the current map from
        the raw parameters, then ONE fused pass (ags_spawn) for back-projection, normals, rejection
        tests, cal_mask, the random voxel filter and the in-place append.  The voxel filter's random
        draw is seeded from torch's CPU generator (torch.manual_seed reproduces it; replicas agree)."""
        dev = self.device
        rgb = dataframe["rgb"].to(dev).float().contiguous()
        depth = dataframe["depth"].to(dev).float().contiguous()
        _, H, W = rgb.shape
        if not self.frames_on_host and not (dataframe["rgb"].is_cuda and dataframe["depth"].is_cuda):
            # keyframes stay resident in HBM: the H2D of a keyframe happens once, here
            # (the reference: mapping/mapper.py:95)
            dataframe = {**dataframe, "rgb": rgb, "depth": depth}
        self.training_data.append(dataframe)
        self.training_performance = torch.cat(
            (self.training_performance, torch.tensor([10.0], device=dev)), 0)
        k = len(self.training_data) - 1
        cam = self._camera(k)
        smooth = O.get_smooth_depth_device(depth)          # (1,H,W); reference: cv2 on the CPU, :297-298
        pred = None
        if self.is_init:
            self._make_contiguous()
            rb = self._render_raw([k], H, W)
            pred = (rb.rgb, rb.depth, rb.opacity)
            self._frame_cost[k] = float(rb.last_instances)      # seeds the cost-balanced frame partition
        n_old = self._means.shape[0]
        st = self._store
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        st.adopt(self, H * W)                              # room for one Gaussian per pixel: no retry needed
        n_new, _, wanted = ops.spawn(rgb, depth, smooth, cam["c2w"], cam["Kinv"], pred, st.buf, n_old, st.cap,
                                     error_thres=self.error_thres, voxel_size=0.02, seed=seed, pool=self._pool)
        assert wanted == n_new
        st.expose(self, n_old + n_new)

    def cal_mask(self, rgb_gt, depth_gt, pred):
        """:470-489 -- spawn where rgb MSE > error_thres, opacity < 0.5 or the render is > 5 % behind."""
        v, _, h, w = rgb_gt.shape
        if pred is None:
            return torch.ones(v * h * w, dtype=torch.bool, device=rgb_gt.device)
        err = torch.mean((rgb_gt - pred["rgb"]) ** 2, dim=1)
        mask = err > self.error_thres
        mask = mask | (pred["opacity"] < 0.5)
        mask = mask | ((depth_gt.squeeze(0) - pred["depth"]) < -0.05 * depth_gt.squeeze(0))
        return mask.reshape(-1)

    # ------------------------------------------------------------------ :491-527 (.th dict format kept)
    def save(self, save_path, index="final"):
        # the map tensors are [:N] views of capacity buffers (or of the symmetric flat buffer of the
        # multi-GPU engine): torch.save would serialise the whole underlying storage, so clone
        c = lambda t: t.detach().clone()
        torch.save({
            "means": c(self._means), "scales": c(self._scales),
            "harmonics": c(self._harmonics), "opacities": c(self._opacities),
            "rotations": c(self._rotations), "view_scores": c(self.view_scores),
            "view_supports": c(self.view_supports), "view_means": c(self.view_means),
            "near": self.scene_near, "far": self.scene_far,
            "use_view_direction": self.use_view_distribution,
            "background_color": self.background_color, "scale_factor": self.scale_factor,
        }, f"{save_path}/map_{index}.th")

    def load(self, model_path):
        s = torch.load(model_path, map_location=self.device, weights_only=False)
        self._means, self._scales, self._harmonics = s["means"], s["scales"], s["harmonics"]
        self._opacities, self._rotations = s["opacities"], s["rotations"]
        self.view_scores, self.view_supports, self.view_means = s["view_scores"], s["view_supports"], s["view_means"]
        self.scene_near, self.scene_far = s["near"], s["far"]
        self.background_color = torch.as_tensor(s["background_color"], dtype=torch.float32).to(self.device)
        self.scale_factor = s["scale_factor"]
        self.is_init = True

    # ------------------------------------------------------------------ :529-590
    @property
    def get_means(self):
        return self._means

    @property
    def get_rotations(self):
        return F.normalize(self._rotations)

    @property
    def get_scales(self):
        return torch.clamp(self.scale_factor * torch.exp(self._scales), min=0, max=0.05)

    @property
    def get_opacities(self):
        return torch.sigmoid(self._opacities)

    @property
    def get_harmonics(self):
        return self._harmonics

    @property
    def get_confidences(self):
        if self.use_view_distribution:
            vv = self.view_means.norm(dim=-1)
            vv = torch.where(torch.isnan(vv), torch.ones_like(vv), vv)
            return torch.clamp(torch.exp(1 - vv) * self.view_scores, min=0, max=1)
        return torch.clamp(1 - 1 / torch.exp(self.view_supports), min=0, max=1)

    @property
    def get_normals(self):
        return F.normalize(O.quaternion_to_matrix(self.get_rotations)[:, :3, 2])

    def get_attr(self):
        return (self.get_means, self.get_harmonics, self.get_opacities, self.get_confidences,
                self.get_scales, self.get_rotations)

    def get_params(self):
        return (self._means, self._harmonics, self._opacities, self._scales, self._rotations)
