"""Time the phases of GaussianMap.update() (spawn / train / post-processing) on a C2-shaped stream of
keyframes.  Development tool (not part of the product path)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from active_gs_b200 import synthetic as syn, operations as O
from active_gs_b200.config import default_gaussian_map_config
from active_gs_b200.gaussian_map import GaussianMap

dev = torch.device("cuda:0")
box, H, W, N = syn.ROOMS[2]
gen = syn.make_room_scene(N, box=box, seed=5)
ext, K = syn.make_cameras(12, box=box, H=H, W=W, seed=6)
src = GaussianMap(default_gaussian_map_config(), dev)
for k, v in gen.items():
    setattr(src, k if k.startswith("view_") else "_" + k, v.to(dev))
gm = GaussianMap(default_gaussian_map_config(), dev)
np.random.seed(0); torch.manual_seed(0)


def sync():
    torch.cuda.synchronize(); return time.time()


for i in range(12):
    with torch.no_grad():
        out = O.GaussianRenderer(ext[i:i + 1].to(dev), K[i:i + 1].to(dev), src.get_attr(), src.background_color,
                                 (0.001, 10.0), (H, W), dev).render_view_all()
    depth = torch.where(out[3][0] > 0.5, out[1][0], torch.full_like(out[1][0], -1.0))
    frame = dict(rgb=out[0][0].clamp(0, 1), depth=depth, extrinsic=ext[i].to(dev), intrinsic=K[i].to(dev),
                 depth_range=torch.tensor([0.0, 5.0]))
    t0 = sync(); gm.add_gaussians(frame)
    t1 = sync(); ctx = gm.begin_training()
    t2 = sync()
    its = []
    for _ in range(10):
        a = sync(); gm.train_step(ctx); its.append(1e3 * (sync() - a))
    if i == 11:
        print("   per-iteration ms (synchronised):", [round(x, 2) for x in its], "instances", ctx.log[-1][2], "visible", ctx.log[-1][3])
    gm.end_training(ctx)
    t3 = sync(); gm.post_processing(); gm.is_init = True
    t4 = sync()
    print(f"kf {i:2d} N={gm._means.shape[0]:7d} spawn {1e3*(t1-t0):7.1f} ms | setup {1e3*(t2-t1):6.1f} | 10 iters {1e3*(t3-t2):6.1f} "
          f"(B={ctx.B}) | post {1e3*(t4-t3):6.1f}")
