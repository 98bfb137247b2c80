"""Where does a GaussianMap.update() go at steady state?  Wall clock per phase WITHOUT intermediate device syncs (the
phases are separated by the syncs the code has anyway), plus a sync-bracketed second pass.  Development tool."""
import sys, os, time, contextlib, io
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from active_gs_b200 import gaussian_map as G

dev = torch.device("cuda:0")
n_upd = 8
state, start, frames, new_frames, cfg, (H, W, N, T) = bench.build_workload(dev, 0, 1, extra=n_upd + 2)
host_new = [{k: (v.pin_memory() if k in ("rgb", "depth") else v) for k, v in f.items()} for f in new_frames]


def run(sync_phases):
    np.random.seed(1234); torch.manual_seed(1234)
    gm = bench.fresh_map(cfg, start, frames, dev, on_host=False, shard=None)
    gm.is_init = True
    acc = {}

    def timed(obj, name, label=None):
        fn = getattr(obj, name)

        def w(*a, **k):
            if sync_phases:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            if sync_phases:
                torch.cuda.synchronize()
            acc[label or name] = acc.get(label or name, 0.0) + time.perf_counter() - t0
            return r
        setattr(obj, name, w)

    for n in ["add_gaussians", "begin_training", "end_training", "post_processing", "train_step"]:
        timed(gm, n)
    from active_gs_b200 import ops, operations as O
    timed(ops, "spawn", "  spawn kernel+readback")
    timed(O, "get_smooth_depth_device", "  bilateral")
    timed(gm, "_render_raw", "  _render_raw (spawn / post)")
    timed(ops, "view_stats_update", "  view_stats_update")
    timed(gm, "_compact", "  _compact")
    with contextlib.redirect_stdout(io.StringIO()):
        for f in host_new[:2]:
            gm.update(f)
        torch.cuda.synchronize(); acc.clear()
        t0 = time.perf_counter()
        per = []
        for f in host_new[2:2 + n_upd]:
            a = time.perf_counter(); gm.update(f); per.append(1e3 * (time.perf_counter() - a))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"sync_phases={sync_phases}: {1e3 * dt / n_upd:.2f} ms per update, N_end={gm._means.shape[0]}, per update {[round(x, 1) for x in per]}")
    for k, v in acc.items():
        print(f"   {k:32s} {1e3 * v / n_upd:7.3f} ms/update")


run(False)
run(True)
