"""Finer timing of add_gaussians() and begin_training() sub-steps (development tool)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import torch.nn.functional as F
from active_gs_b200 import synthetic as syn, operations as O, lib as L
from active_gs_b200.config import default_gaussian_map_config
from active_gs_b200.gaussian_map import GaussianMap, _TrainEngine, WeightedSampler

dev = torch.device("cuda:0")
box, H, W, N = syn.ROOMS[2]
gen = syn.make_room_scene(N, box=box, seed=5)
ext, K = syn.make_cameras(9, box=box, H=H, W=W, seed=6)
gm = GaussianMap(default_gaussian_map_config(), dev)
for k, v in gen.items():
    setattr(gm, k if k.startswith("view_") else "_" + k, v.to(dev))
gm.is_init = True
frames = []
with torch.no_grad():
    out = O.GaussianRenderer(ext.to(dev), K.to(dev), gm.get_attr(), gm.background_color, (0.001, 10.0), (H, W), dev).render_view_all()
for i in range(9):
    frames.append(dict(rgb=out[0][i].clamp(0, 1), depth=out[1][i], extrinsic=ext[i].to(dev), intrinsic=K[i].to(dev),
                       depth_range=torch.tensor([0.0, 5.0])))
gm.training_data = frames[:8]
gm.training_performance = torch.full((8,), 10.0, device=dev)
T = lambda: (torch.cuda.synchronize(), time.time())[1]
fr = frames[8]
for rep in range(3):
    rgb, depth = fr["rgb"], fr["depth"]; intrinsic, extrinsic = fr["intrinsic"], fr["extrinsic"]
    t = [T()]
    d_np = depth.squeeze(0).cpu().numpy(); t.append(T())
    sm = O.get_smooth_depth(d_np); t.append(T())
    smooth = torch.tensor(sm, device=dev).unsqueeze(0); t.append(T())
    origins, directions = O.get_world_rays(H, W, extrinsic, intrinsic, dev); pcd = origins + directions * depth.view(-1, 1); t.append(T())
    n_cam = O.depth2normal(smooth, (depth > 0).view(1, H, W), fov=(np.pi / 3, np.pi / 3)); t.append(T())
    r = O.GaussianRenderer(extrinsic[None], intrinsic[None], gm.get_attr(), gm.background_color, (0.001, 10.0), (H, W), dev).render_view_all(); t.append(T())
    nrm = F.normalize(torch.randn(H * W, 3, device=dev)); rot, _ = O.normal2rotation(nrm); t.append(T())
    sel = torch.rand(H * W, device=dev) > 0.7
    keep = O.voxel_downsample(pcd[sel]); t.append(T())
    cat = torch.cat((gm._means, pcd[sel][keep])); t.append(T())
    names = ["D2H depth", "cv2 bilateral", "H2D", "world rays", "depth2normal", "render map", "normal2rotation", "voxel_downsample", "one cat"]
    print("spawn parts ms:", {n: round(1e3 * (b - a), 2) for n, a, b in zip(names, t[:-1], t[1:])})
for rep in range(3):
    t = [T()]
    gm._make_contiguous(); t.append(T())
    rows = gm._camera_rows(range(len(gm.training_data))); t.append(T())
    eng = _TrainEngine(gm, 8, H, W, None, cam_table=rows.to(dev)); t.append(T())
    print("setup parts ms:", {n: round(1e3 * (b - a), 2) for n, a, b in zip(["contiguous", "camera table", "engine init"], t[:-1], t[1:])})
    del eng
