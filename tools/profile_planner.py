"""Time the planners' utility evaluation: 100 candidate views at 128x128
(render_ratio 0.25 of 512x512, config/planner/confidence.yaml) over the BASELINE config[1] map and a
0.2 m voxel grid, through active_gs_b200.planning.Confidence.cal_utility.  Development tool."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from types import SimpleNamespace as ns
import numpy as np, torch
from active_gs_b200 import synthetic as syn, planning
from active_gs_b200.config import default_gaussian_map_config
from active_gs_b200.gaussian_map import GaussianMap

dev = torch.device("cuda:0")
box, H, W, N = syn.ROOMS[2]
state = syn.make_room_scene(N, box=box, seed=1002)
gm = GaussianMap(default_gaussian_map_config(), dev)
for k, v in state.items():
    setattr(gm, k if k.startswith("view_") else "_" + k, v.to(dev))
ext, K = syn.make_cameras(100, box=box, H=128, W=128, hfov=60.0, seed=77)
dim = [int(np.ceil(b / 0.2)) for b in box]
g = torch.meshgrid(*[torch.arange(d) for d in dim], indexing="ij")
centers = torch.stack([(g[k].float() + 0.5) * (box[k] / dim[k]) for k in range(3)], -1).reshape(-1, 3)
gen = torch.Generator().manual_seed(0)
vm = ns(voxel_centers=centers, unexplored_mask=torch.rand(centers.shape[0], generator=gen) < 0.5)
sim = ns(resolution=np.array([512, 512]), depth_range=[0.0, 5.0], intrinsic=K[0], has_missing_surface=False)
cfg = ns(render_ratio=0.25, explore_weight=1000.0)
pl = planning.Confidence(cfg, dev)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    u, t_util = pl.cal_utility(gm, vm, ext, sim)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: cal_utility of 100 views x 128x128, N={N}, M={centers.shape[0]} voxels: {1e3*dt:.2f} ms "
          f"(t_utility {1e3*t_util:.2f} ms), best view {int(u.argmax())}, utility range {float(u.min()):.2f}..{float(u.max()):.2f}")
