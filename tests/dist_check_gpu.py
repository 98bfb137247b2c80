"""Multi-GPU check (run under torchrun, one process per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check_gpu.py

Trains the same small map for a few iterations with
  (a) one GPU's worth of frames on a single rank replica (reference: no sharding, rank 0 only),
  (b) frame sharding + NCCL all-reduce + replicated Adam,
  (c) frame sharding + the fused NVLink kernel, peer-pointer path,
  (d) frame sharding + the fused NVLink kernel, NVLS multimem path (if the fabric supports it),
and checks that all ranks hold identical parameters and that (b), (c), (d) agree with (a) -- once with a
keyframe batch that divides by the world size and once with one that does not (padded slots with loss
weight 0, the normal case for the first keyframes of a mission).
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from active_gs_b200 import synthetic as syn, operations as O  # noqa: E402
from active_gs_b200.config import default_gaussian_map_config  # noqa: E402
from active_gs_b200.gaussian_map import GaussianMap  # noqa: E402
from active_gs_b200.distributed import FrameShard  # noqa: E402

NAMES = ["_means", "_scales", "_rotations", "_opacities", "_harmonics"]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    box, H, W, N = (6.0, 4.5, 2.7), 120, 160, 30000
    T = 4 * world
    gen = syn.make_room_scene(N, box=box, seed=11)
    gen["scales"][:, :2] += 1.0
    ext, K = syn.make_cameras(T, box=box, H=H, W=W, seed=12)
    cfg = default_gaussian_map_config()
    cfg.sampler.batch_size = T
    src = GaussianMap(cfg, dev)
    for k, v in gen.items():
        setattr(src, k if k.startswith("view_") else "_" + k, v.to(dev))
    frames = []
    with torch.no_grad():
        out = O.GaussianRenderer(ext.to(dev), K.to(dev), src.get_attr(), src.background_color, (0.001, 10.0),
                                 (H, W), dev).render_view_all()
    for i in range(T):
        frames.append(dict(rgb=out[0][i].clamp(0, 1), depth=out[1][i], extrinsic=ext[i], intrinsic=K[i],
                           depth_range=torch.tensor([0.0, 5.0])))
    start = syn.perturb_state(gen, seed=13)

    def run(shard, steps=4, batch=None):
        cfg.sampler.batch_size = T if batch is None else batch
        gm = GaussianMap(cfg, dev)
        for k, v in start.items():
            setattr(gm, k if k.startswith("view_") else "_" + k, v.clone().to(dev))
        gm.training_data = frames
        gm.training_performance = torch.full((T,), 10.0, device=dev)
        gm.dist = shard
        np.random.seed(5)
        ctx = gm.begin_training()
        losses = [gm.train_step(ctx) for _ in range(steps)]
        gm.end_training(ctx)
        torch.cuda.synchronize()
        return [getattr(gm, n).detach().clone() for n in NAMES], losses, gm.training_performance.clone()

    ok = True
    mc = False
    for batch in (T, T - 1):
        ref, ref_losses, ref_perf = run(None, batch=batch)              # every rank: the whole batch on its own GPU
        results = {"nccl": run(FrameShard(fused=False), batch=batch)}
        sh = FrameShard(fused=True); sh.use_multicast = False
        results["fused-peer"] = run(sh, batch=batch)
        sh2 = FrameShard(fused=True); sh2.use_multicast = True
        results["fused-multimem"] = run(sh2, batch=batch)
        mc = sh2._flat is not None and sh2._flat.grad_mc != 0
        for name, (params, losses, perf) in results.items():
            # (1) replicas identical across ranks
            for p in params:
                lo, hi = p.clone(), p.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                same = bool(torch.equal(torch.nan_to_num(lo), torch.nan_to_num(hi)))
                ok &= same
                if not same and rank == 0:
                    print(f"[{name}] replicas differ: max {(hi - lo).abs().max().item():.3e}")
            # (2) agrees with the unsharded run
            errs = [((a - b).norm() / b.norm().clamp_min(1e-30)).item() for a, b in zip(params, ref) if True]
            errs[1] = ((params[1][:, :2] - ref[1][:, :2]).norm() / ref[1][:, :2].norm()).item()   # skip the -1e10 lane
            lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
            perr = (perf - ref_perf).abs().max().item()
            good = max(errs) < 2e-4 and lerr < 1e-4 and perr < 1e-5
            ok &= good
            if rank == 0:
                print(f"[batch {batch}{' (padded)' if batch % world else ''}] "
                      f"[{name}{' (NVLS multimem active)' if name == 'fused-multimem' and mc else ''}] "
                      f"param l2 err vs unsharded {['%.1e' % e for e in errs]} loss err {lerr:.1e} perf err {perr:.1e} "
                      f"{'ok' if good else 'FAIL'}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK", "PASS" if int(flag) else "FAIL", f"(world {world}, multicast {'yes' if mc else 'no'})")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
