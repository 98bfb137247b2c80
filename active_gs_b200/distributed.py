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


class FrameShard:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def local_batch(self, B_global):
        if B_global % self.world != 0:
            raise ValueError(f"global keyframe batch {B_global} must divide by world size {self.world}")
        return B_global // self.world

    def my_frames(self, ids):
        """contiguous slice of the sampled ids owned by this rank"""
        ids = np.asarray(ids)
        b = len(ids) // self.world
        return ids[self.rank * b:(self.rank + 1) * b]

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
