"""Device timeline of the end-to-end step: CUDA events on the compute and copy streams."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from active_gs_b200 import gaussian_map as G, lib as L, ops

dev = torch.device("cuda:0")
state, start, frames, cfg, (H, W, N, T) = bench.build_workload(dev, 0, 1)
np.random.seed(1234)
gm = bench.fresh_map(cfg, start, frames, dev, on_host=True, shard=None)
ctx = gm.begin_training()
eng = ctx.eng
lib = L.load()
marks = []
def ev(stream=None):
    e = torch.cuda.Event(enable_timing=True); e.record(stream or torch.cuda.current_stream()); return e
def wrap(name):
    f = getattr(lib, name)
    def w(*a):
        marks.append((name + ":b", ev())); r = f(*a); marks.append((name + ":e", ev())); return r
    setattr(lib, name, w)
for n in ["ags_render_forward", "ags_loss_forward_backward", "ags_render_backward", "ags_adam_step"]:
    wrap(n)
osb, opf = eng.set_batch, eng.prefetch_next
def sb(*a, **k):
    marks.append(("copy:b", ev(eng.copy_stream))); r = osb(*a, **k); marks.append(("copy:e", ev(eng.copy_stream))); return r
def pf(*a, **k):
    marks.append(("pref:b", ev(eng.copy_stream))); r = opf(*a, **k); marks.append(("pref:e", ev(eng.copy_stream))); return r
eng.set_batch, eng.prefetch_next = sb, pf
for _ in range(5):
    gm.train_step(ctx)
torch.cuda.synchronize(); marks.clear()
for _ in range(6):
    gm.train_step(ctx)
torch.cuda.synchronize()
t0 = marks[0][1]
for n, e in marks:
    print(f"{t0.elapsed_time(e):9.3f} ms  {n}")
