"""Frame sharding of the keyframe batch across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  Every rank holds
the full SoA parameter set and Adam state; within an iteration rank r renders its slice of the
sampled keyframe batch.  Exchanges per iteration:
  1. all-reduce(SUM) of the per-pixel visibility count (H,W) int32 -- quirk Q1 couples the frames'
     consistency terms through the sum of their visibility masks (mapping/gaussian_map.py:116-117)
  2. all-reduce(SUM) of the raw-parameter gradients, 56 B per Gaussian, as ONE flat buffer
  3. all-gather of the per-frame performance scalars (sampler weights, mapping/utils.py:206-218)
The sampler draw uses numpy's global RNG with the same seed on every rank, so the ids agree without
a broadcast.  Loss normalisers use the global batch size (AgsLossArgs.B_total).
"""
import numpy as np
import torch
import torch.distributed as dist


class SymmetricFlat:
    """Symmetric (peer-mapped) flat buffers for the fused exchange: parameters, gradients, both of
    `numel_padded` floats at the same offset on every GPU of the group, plus their NVLS multicast
    addresses when the fabric supports it."""

    def __init__(self, group, numel, device):
        import torch.distributed._symmetric_memory as symm
        world = dist.get_world_size(group)
        q = 4 * world
        self.numel = numel
        self.numel_padded = (numel + q - 1) // q * q
        self.param = symm.empty(self.numel_padded, dtype=torch.float32, device=device)
        self.grad = symm.empty(self.numel_padded, dtype=torch.float32, device=device)
        self.stats = symm.empty(72, dtype=torch.int32, device=device)           # AGS_NUM_STATS      # AgsRenderArgs.stats, peer readable
        self.param.zero_(); self.grad.zero_(); self.stats.zero_()
        g = group if group is not None else dist.group.WORLD
        self.h_param = symm.rendezvous(self.param, g)
        self.h_grad = symm.rendezvous(self.grad, g)
        self.h_stats = symm.rendezvous(self.stats, g)
        self.param_ptrs = [int(p) for p in self.h_param.buffer_ptrs]
        self.grad_ptrs = [int(p) for p in self.h_grad.buffer_ptrs]
        self.stats_ptrs = [int(p) for p in self.h_stats.buffer_ptrs]
        mc = bool(self.h_param.has_multicast_support) if hasattr(self.h_param, "has_multicast_support") else False
        self.param_mc = int(self.h_param.multicast_ptr) if mc else 0
        self.grad_mc = int(self.h_grad.multicast_ptr) if mc else 0

    def barrier(self):
        """cross-GPU barrier on the current stream (device side, no host sync)"""
        self.h_grad.barrier()


class SymmetricAux:
    """Symmetric buffers of the two small per-iteration exchanges (csrc/dist_loss.cu): the (H*W) int32
    visibility-count plane of quirk Q1 and the (world*nterm) float gather buffer of the loss terms."""

    def __init__(self, group, P, nterm, device):
        import torch.distributed._symmetric_memory as symm
        world = dist.get_world_size(group)
        self.P, self.nterm = P, nterm
        self.vis = symm.empty(P, dtype=torch.int32, device=device)
        self.gather = symm.empty(2 * world * nterm, dtype=torch.float32, device=device)     # two halves: iteration parity
        self.vis.zero_(); self.gather.zero_()
        g = group if group is not None else dist.group.WORLD
        self.h_vis = symm.rendezvous(self.vis, g)
        self.h_gather = symm.rendezvous(self.gather, g)
        self.vis_ptrs = [int(p) for p in self.h_vis.buffer_ptrs]
        self.gather_ptrs = [int(p) for p in self.h_gather.buffer_ptrs]
        mc = bool(getattr(self.h_vis, "has_multicast_support", False))
        self.vis_mc = int(self.h_vis.multicast_ptr) if mc else 0
        self.gather_mc = int(self.h_gather.multicast_ptr) if mc else 0

    def barrier(self):
        self.h_vis.barrier()


class SymmetricSync:
    """The flags array of the folded cross-GPU ordering (AgsDistSync, include/ags_b200.h): AGS_SYNC_WORDS int32 per
    rank, peer-mapped; producers store the iteration's epoch into every peer's array, consumers spin on their own."""

    def __init__(self, group, device, words=64):
        import torch.distributed._symmetric_memory as symm
        self.flags = symm.empty(words, dtype=torch.int32, device=device)
        self.flags.zero_()
        g = group if group is not None else dist.group.WORLD
        self.h = symm.rendezvous(self.flags, g)
        self.ptrs = [int(p) for p in self.h.buffer_ptrs]
        self.epoch = 0
        torch.cuda.synchronize(device)
        self.h.barrier()                       # everybody's flags are zero before anybody signals


class FrameShard:
    FLAT_RESERVE = 14 * (1 << 20)            # floats of head-room in the symmetric parameter / gradient buffers

    def __init__(self, group=None, fused=False):
        self.flat_allocations = 0
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # fused = reduce-scatter -> Adam -> all-gather in ONE kernel over NVLink peer memory
        # (csrc/dist_adam.cu) instead of NCCL all-reduce + replicated Adam
        self.fused = fused
        # NVLS multimem for the gradient/parameter exchange: measured faster from 4 GPUs up (N=8: 77-110 us
        # vs 125 us with peer loads), slower at N=2 (57 vs 43 us)
        self.use_multicast = self.world >= 4
        self._flat = None
        self._aux = None
        self._sync = None
        # fold the four per-iteration barriers into the exchange kernels (signal / wait on symmetric flags)
        # instead of separate barrier launches; AGS_DIST_BARRIER=1 keeps the barrier launches (A/B runs)
        import os
        self.folded = os.environ.get("AGS_DIST_BARRIER", "0") != "1"
        self._nccl_warm = False

    def warm_up(self, device):
        """NCCL connects its channels lazily, on the first collective of each kind and size class: measured 259 ms
        for the first all-reduce of the sharded prune pass (one per mission, but it would land inside a keyframe
        update).  Pay it when the shard is set up: one small and one Gaussian-count-sized int32 all-reduce."""
        if self._nccl_warm or dist.get_backend(self.group) != "nccl":
            return
        for n in (1, 1 << 20):
            t = torch.zeros(n, dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize(device)
        self._nccl_warm = True

    def flat_buffers(self, numel, device):
        """symmetric buffers with head-room, re-allocated (collective!) only when the map outgrows them"""
        if self._flat is None or self._flat.numel_padded < numel:
            # allocation + rendezvous of symmetric memory costs ~100 ms per buffer on an 8-GPU box (measured:
            # one re-allocation inside a 2-update timed region = 20 ms/step), the bytes are cheap: 2x plus room
            # for ~1M further Gaussians, so a mapping run re-allocates O(log) times
            self._flat = SymmetricFlat(self.group, 2 * int(numel) + self.FLAT_RESERVE, device)
            self.flat_allocations += 1
        return self._flat

    def sync_buffers(self, device):
        if self._sync is None:
            self._sync = SymmetricSync(self.group, device)
        return self._sync

    def aux_buffers(self, P, nterm, device):
        """symmetric buffers of the visibility / loss-term exchange (collective allocation)"""
        if self._aux is None or self._aux.P != P or self._aux.nterm != nterm:
            self._aux = SymmetricAux(self.group, P, nterm, device)
        return self._aux

    def all_reduce_max_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def local_batch(self, B_global):
        """frames per rank: the sampled batch is padded to a multiple of the world size (the reference
        sampler yields T < 8 frames for the first keyframes, mapping/utils.py:196-204); padded slots are
        rendered from a repeated keyframe with loss weight 0, so every rank enters every exchange"""
        return max(1, -(-int(B_global) // self.world))

    def pad(self, ids):
        """ids -> array of local_batch(len(ids)) * world entries, -1 = padding"""
        ids = [int(i) for i in ids]
        total = self.local_batch(len(ids)) * self.world
        return np.asarray(ids + [-1] * (total - len(ids)))

    def my_frames(self, ids):
        """contiguous slice of the (padded, balanced) sampled ids owned by this rank"""
        ids = np.asarray(ids)
        b = len(ids) // self.world
        return ids[self.rank * b:(self.rank + 1) * b]

    def pinned_slot(self, j):
        """(rank, local slot) of the j-th always-selected (active) keyframe: round robin over the
        ranks at fixed slots, so every rank can prefetch its share before the sampler draw"""
        return j % self.world, j // self.world

    def balance(self, ids, n_active, cost):
        """Partition the sampled keyframe ids over the ranks (the gradient sum does not depend on the
        partition) so that the slowest rank is as fast as possible: the first `n_active` ids stay at
        their pinned slots, the others are placed longest-first on the least loaded rank with a free
        slot (LPT with a cardinality constraint).  `cost`: id -> instances of its last render (missing
        ids count as the mean).  Returns the ids reordered so that rank r owns [r*b, (r+1)*b), b =
        local_batch(len(ids)); slots left over when len(ids) is not a multiple of the world size hold -1.
        Deterministic: every rank computes the same answer from the same gathered costs."""
        ids = [int(i) for i in ids]
        n = len(ids)
        W, b = self.world, self.local_batch(n)
        get = cost.get
        c = [get(i) for i in ids]
        if None in c:
            known = [x for x in c if x is not None]
            mean = (sum(known) / len(known)) if known else 0.0
            c = [mean if x is None else float(x) for x in c]
        slots = [[] for _ in range(W)]                 # pinned ids first (slot j // W of rank j % W), then LPT
        load = [0.0] * W
        na = min(n_active, n)
        for j in range(na):
            r = j % W
            slots[r].append(ids[j])
            load[r] += c[j]
        # longest first onto the least loaded rank that still has a free slot (ties: lower rank): a heap of
        # (load, rank); a rank leaves the heap when it is full
        import heapq
        heap = [(load[r], r) for r in range(W) if len(slots[r]) < b]
        heapq.heapify(heap)
        for j in sorted(range(na, n), key=lambda j: (-c[j], j)):
            l, r = heap[0]
            sr = slots[r]
            sr.append(ids[j])
            if len(sr) < b:
                heapq.heapreplace(heap, (l + c[j], r))
            else:
                heapq.heappop(heap)
        out = np.full(W * b, -1, dtype=np.int64)       # -1 = padding
        for r in range(W):
            sr = slots[r]
            out[r * b:r * b + len(sr)] = sr
        return out

    def all_reduce_sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_grads_(self, grads):
        """`grads` are views into one flat buffer (engine allocates them that way): one collective."""
        base = grads[0]._base if grads[0]._base is not None else None
        if base is not None and all(g._base is base for g in grads):
            dist.all_reduce(base, op=dist.ReduceOp.SUM, group=self.group)
        else:
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)

    def all_gather_into_(self, out, local):
        """out (world*n,) <- concatenation over ranks of local (n,), on the current stream"""
        dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def gather_perf(self, perf_local, ids):
        """per-frame performance of the whole batch, ordered like `ids` (host tensors)"""
        out = [torch.empty_like(perf_local) for _ in range(self.world)]
        if perf_local.is_cuda or dist.get_backend(self.group) == "gloo":
            dist.all_gather(out, perf_local, group=self.group)
        else:
            dev = torch.device("cuda", torch.cuda.current_device())
            outd = [torch.empty_like(perf_local, device=dev) for _ in range(self.world)]
            dist.all_gather(outd, perf_local.to(dev), group=self.group)
            out = [o.cpu() for o in outd]
        return torch.cat(out)
