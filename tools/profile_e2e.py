"""Host/device timeline of the end-to-end training step (keyframes in pinned host memory):
where does the time between two steps go?  Development tool."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from active_gs_b200 import gaussian_map as G

dev = torch.device("cuda:0")
state, start, frames, cfg, (H, W, N, T) = bench.build_workload(dev, 0, 1)
for on_host in (False, True):
    np.random.seed(1234)
    gm = bench.fresh_map(cfg, start, frames, dev, on_host=on_host, shard=None)
    ctx = gm.begin_training()
    eng = ctx.eng
    acc = {}
    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
        return w
    eng.set_batch = timed("set_batch", eng.set_batch)
    eng.iterate = timed("iterate", eng.iterate)
    eng.prefetch_next = timed("prefetch", eng.prefetch_next)
    eng.fetch = timed("fetch(wait)", eng.fetch)
    ctx.sampler.next_ids = timed("sampler", ctx.sampler.next_ids)
    for _ in range(5):
        gm.train_step(ctx)
    torch.cuda.synchronize(); acc.clear()
    # copy duration on the copy stream
    K = 200
    t0 = time.perf_counter()
    for _ in range(K):
        gm.train_step(ctx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"on_host={on_host}: {1e3*dt/K:.3f} ms/step; host ms/step:", {k: round(1e3 * v / K, 3) for k, v in acc.items()},
          "other:", round(1e3 * (dt - sum(acc.values())) / K, 3))
# raw copy pattern as set_batch issues it: 8 frames x (rgb, depth) into (B,3,H,W)/(B,1,H,W) slices
rgb_gt, depth_gt = eng.gt[0]
st = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
td = gm.training_data
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        e0.record(st)
        for k in range(8):
            rgb_gt[k].copy_(td[k]["rgb"], non_blocking=True)
            depth_gt[k].copy_(td[k]["depth"], non_blocking=True)
        e1.record(st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"16 copies: host issue {1e3*(t1-t0):.3f} ms, device {e0.elapsed_time(e1):.3f} ms; pinned={td[0]['rgb'].is_pinned()} contiguous={td[0]['rgb'].is_contiguous()}")
