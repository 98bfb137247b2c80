"""torchrun tool: where does GaussianMap.update() go in the fused multi-GPU mode?  Wall clock per phase on rank 0
(sync-bracketed).  Development tool."""
import sys, os, time, contextlib, io
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch, torch.distributed as dist
import bench
from active_gs_b200 import gaussian_map as G, ops, operations as O
from active_gs_b200.distributed import FrameShard

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
shard = FrameShard(fused=True)
n_upd = 4
state, start, frames, new_frames, cfg, (H, W, N, T) = bench.build_workload(dev, rank, world, extra=n_upd + 2)
host_new = [{k: (v.pin_memory() if k in ("rgb", "depth") else v) for k, v in f.items()} for f in new_frames]
np.random.seed(1234); torch.manual_seed(1234)
gm = bench.fresh_map(cfg, start, frames, dev, on_host=False, shard=shard)
gm.is_init = True
acc = {}


def timed(obj, name, label=None):
    fn = getattr(obj, name)

    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc.setdefault(label or name, []).append(1e3 * (time.perf_counter() - t0))
        return r
    setattr(obj, name, w)


for n in ["add_gaussians", "begin_training", "end_training", "post_processing", "train_step", "_render_raw", "_compact"]:
    timed(gm, n)
timed(gm._store, "adopt", "  store.adopt")
timed(shard, "flat_buffers", "  flat_buffers")
timed(shard, "all_reduce_sum_", "  all_reduce_sum_")
timed(ops, "spawn", "  spawn")
with contextlib.redirect_stdout(io.StringIO()):
    for i, f in enumerate(host_new[:2 + n_upd]):
        if i == 2:
            acc.clear()
        t0 = time.perf_counter(); gm.update(f); torch.cuda.synchronize()
        acc.setdefault("UPDATE", []).append(1e3 * (time.perf_counter() - t0))
if rank == 0:
    for k, v in acc.items():
        print(f"{k:24s} n={len(v):3d} total {sum(v):8.2f} ms  max {max(v):8.2f}  each {[round(x, 2) for x in v[:12]]}", flush=True)
dist.destroy_process_group()
