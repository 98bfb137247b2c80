"""Generate tests/golden/spawn_golden.pt by EXECUTING THE REFERENCE'S OWN PYTHON (build container
only; the output is committed because /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_spawn.py

Pinned here (rows a13-a15 and f1 of SURVEY.md section 8):
  * GaussianMap.add_gaussians (mapping/gaussian_map.py:294-468): first keyframe (no map yet) and a
    keyframe added to an initialised map (cal_mask :470-489 uses a render of the map).  The random
    voxel filter is replaced by "keep everything" for the candidate fixtures (its own output depends
    on torch.randperm) and run for real once, for the occupancy / one-per-voxel properties.
  * get_smooth_depth (utils/operations.py:161-169, OpenCV bilateral filter).
  * VoxelMap.cal_visible_mask (mapping/voxel_map.py:226-278) and the per-view utility arithmetic of
    planning/confidence.py:69-101 / planning/exploration.py:62-86, by running
    Confidence.cal_utility / Exploration.cal_utility on a small voxel map.
As in make_golden.py the absent native rasterizer is satisfied by oracle/rasterizer_ref.py.
"""
import os
import sys
import types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts REPO and /root/reference on sys.path)
from active_gs_b200 import synthetic as syn  # noqa: E402


def keyframe(gmap, ops, gen_state, ext, K, hw, seed, holes=True):
    gm = gmap.GaussianMap(mg.cfg_namespace(), "cpu")
    mg.load_state(gm, gen_state)
    with torch.no_grad():
        rgb, depth, _, opacity, *_ = ops.GaussianRenderer(ext[None], K[None], gm.get_attr(), gm.background_color,
                                                          (0.001, 10.0), hw, "cpu").render_view_all()
    d = torch.where(opacity[0] > 0.5, depth[0], torch.full_like(depth[0], -1.0))
    if holes:
        g = torch.Generator().manual_seed(seed)
        d[torch.rand(d.shape, generator=g) < 0.03] = -1.0
    return dict(rgb=rgb[0].clamp(0, 1), depth=d, extrinsic=ext, intrinsic=K, depth_range=torch.tensor([0.0, 5.0]))


def run_add(gmap, frame, state=None, keep_all=True, seed=0):
    gm = gmap.GaussianMap(mg.cfg_namespace(), "cpu")
    if state is not None:
        mg.load_state(gm, state)
        gm.is_init = True
        gm.training_performance = torch.zeros(0)
    captured = {}
    real = gmap.voxel_downsample

    def keep_everything(points, *a, **k):
        captured["points"] = points.clone()
        return torch.arange(points.shape[0])

    def real_logged(points, *a, **k):
        captured["points"] = points.clone()
        out = real(points, *a, **k)
        captured["selected"] = out.clone()
        return out

    gmap.voxel_downsample = keep_everything if keep_all else real_logged
    n0 = gm._means.shape[0]
    torch.manual_seed(seed)
    try:
        gm.add_gaussians(frame)
    finally:
        gmap.voxel_downsample = real
    new = {k: v[n0:].clone() for k, v in mg.dump_state(gm).items()}
    return new, captured, gm


def main():
    ops, mutils, gmap = mg.import_reference()
    G = {}
    box = (3.0, 2.5, 2.0)
    hw = (40, 56)
    gen = syn.make_room_scene(6000, box=box, seed=91, furniture=3)
    gen["scales"][:, :2] += 1.6                               # disks large enough to close a 56x40 image
    ext, K = syn.make_cameras(3, box=box, H=hw[0], W=hw[1], hfov=70.0, seed=92)

    # ---- first keyframe: no map yet -> every valid pixel is a candidate
    f0 = keyframe(gmap, ops, gen, ext[0], K[0], hw, seed=1)
    G["smooth_depth"] = dict(depth=f0["depth"].clone(),
                             out=torch.tensor(ops.get_smooth_depth(f0["depth"].squeeze(0).numpy())))
    new0, cap0, _ = run_add(gmap, f0, state=None, keep_all=True)
    G["first"] = dict(frame=f0, new=new0, hw=hw)

    # ---- second keyframe on an initialised map: a thinned + recoloured copy of the generating scene,
    # so that the three cal_mask criteria (rgb error, opacity < 0.5, render behind the sensor) all fire
    g = torch.Generator().manual_seed(7)
    keep = torch.rand(6000, generator=g) < 0.55
    state = {k: v[keep].clone() for k, v in gen.items()}
    far = state["means"][:, 0] > 0.3 * box[0]
    state["harmonics"][far] = (state["harmonics"][far] + 0.7).clamp(0, 1.6)      # rgb error > thres there
    push = (state["means"][:, 1] > 0.2 * box[1]) & ~far
    state["means"][push] *= 1.12                                                   # rendered behind the sensor depth
    f1 = keyframe(gmap, ops, gen, ext[1], K[1], hw, seed=2)
    new1, cap1, _ = run_add(gmap, f1, state=state, keep_all=True)
    G["second"] = dict(frame=f1, state=state, new=new1, hw=hw)

    # ---- the real voxel filter once (occupancy / one-per-voxel properties; the choice is random)
    new2, cap2, _ = run_add(gmap, f0, state=None, keep_all=False, seed=5)
    G["voxel"] = dict(points=cap2["points"], selected=cap2["selected"], n_new=new2["means"].shape[0], voxel_size=0.02)
    coarse = cap2["points"] * 0.25                                                # same cloud, effectively 8 cm voxels
    torch.manual_seed(6)
    G["voxel_coarse"] = dict(points=coarse, selected=ops.voxel_downsample(coarse), voxel_size=0.02)

    # ---- planner utilities on a small voxel map
    for n in ["utils.common"]:
        m = sys.modules[n]
        m.Planner2Gui = mg._Anything
    import mapping.voxel_map as vmod
    import planning.confidence as pconf
    import planning.exploration as pexp
    ns = types.SimpleNamespace
    vcfg = ns(min_gaussian_per_voxel=5, map_resolution=[0.25, 0.25, 0.25], safety_margin=0.3)
    bbox = [[-0.5 * box[0], -0.5 * box[1], 0.0], [0.5 * box[0], 0.5 * box[1], box[2]]]
    lo = gen["means"].min(0).values - 0.1
    hi = gen["means"].max(0).values + 0.1
    bbox = [lo.tolist(), hi.tolist()]
    vm = vmod.VoxelMap(vcfg, bbox, "cpu")
    g = torch.Generator().manual_seed(11)
    vm.unexplored_mask = torch.rand(vm.voxel_centers.shape[0], generator=g) < 0.6
    pcfg = ns(pitch_angle=None, robot_size=0.3, radius=0.5, init_pose=torch.eye(4).tolist(), path_length_factor=0.5,
              use_confidence=True, sample_num=6, max_roi_sample_num=0, render_ratio=0.25, explore_weight=2.0)
    cand_ext, cand_K = syn.make_cameras(6, box=box, H=128, W=128, hfov=60.0, seed=93)
    sim = ns(resolution=np.array([128, 128]), depth_range=[0.0, 2.5], intrinsic=cand_K[0], has_missing_surface=False)
    gm = gmap.GaussianMap(mg.cfg_namespace(), "cpu")
    mg.load_state(gm, state)
    gm.view_scores = torch.rand(state["means"].shape[0], generator=g) * 1.5
    gm.view_means = torch.nn.functional.normalize(torch.randn(state["means"].shape[0], 3, generator=g), dim=-1) * \
        torch.rand(state["means"].shape[0], 1, generator=g)
    planner = pconf.Confidence(pcfg, "cpu")
    util, _ = planner.cal_utility(gm, vm, cand_ext, sim)
    explorer = pexp.Exploration(pcfg, "cpu")
    util_e, _ = explorer.cal_utility(gm, vm, cand_ext, sim)
    # per-view pieces for the restated arithmetic
    r = ops.GaussianRenderer(cand_ext, cand_K, gm.get_attr(), gm.background_color, (0.001, 10.0), (32, 32), "cpu")
    vis = []
    for i in range(6):
        out = r.render_view(i)
        dv = out[1][0].clone()
        dv[dv < 0.001] = 10000
        dv = torch.clamp(dv, min=sim.depth_range[0], max=sim.depth_range[1])
        vis.append(vm.cal_visible_mask(cand_ext[i], cand_K[i], dv))
    # ---- VoxelMap.update_utility (mapping/voxel_map.py:62-116): voxels with many opaque low-confidence surfels
    vcfg2 = ns(min_gaussian_per_voxel=3, map_resolution=[0.25, 0.25, 0.25], safety_margin=0.3)
    vm2 = vmod.VoxelMap(vcfg2, bbox, "cpu")
    gm.view_scores = gm.view_scores * 0.25                     # most surfels below the 0.3 confidence threshold
    vm2.update_utility(gm, use_confidence=True)
    G["voxel_roi"] = dict(state=mg.dump_state(gm), bbox=vm2.bbox.clone(), size=vm2.size.clone(), dim=vm2.dim.clone(),
                          min_gaussian_per_voxel=3, confidence_thres=0.3, voxel_normal=vm2.voxel_normal.clone())
    gm.view_scores = gm.view_scores * 4.0
    map_state = mg.dump_state(gm)
    G["planner"] = dict(state=map_state, ext=cand_ext, K=cand_K, hw=(32, 32), depth_range=sim.depth_range,
                        voxel_centers=vm.voxel_centers.clone(), unexplored=vm.unexplored_mask.clone(),
                        explore_weight=pcfg.explore_weight, utility_confidence=util.clone(),
                        utility_exploration=util_e.clone(), visible=torch.stack(vis))

    out = os.path.join(HERE, "spawn_golden.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
    print("first: candidates", new0["means"].shape[0], "of", hw[0] * hw[1], "| second:", new1["means"].shape[0],
          "| voxel filter kept", G["voxel"]["n_new"], "of", cap2["points"].shape[0],
          "| coarse kept", G["voxel_coarse"]["selected"].numel())
    print("voxel roi: voxels flagged", int((G["voxel_roi"]["voxel_normal"].norm(dim=1) > 0).sum()), "of", G["voxel_roi"]["voxel_normal"].shape[0])
    print("utility (confidence)", util, "\nutility (exploration)", util_e, "visible voxels", G["planner"]["visible"].sum(1))


if __name__ == "__main__":
    main()
