"""Multi-GPU correctness inside pytest: spawns tests/dist_check_gpu.py under torchrun at every world
size the box offers (2, 4, 8) and requires its PASS line -- frame sharding with the NCCL baseline, the
fused NVLink kernels (peer pointers and NVLS multimem), the cost-balanced partition and the padded
(not divisible) keyframe batch, each against the unsharded run.  Skipped on boxes with one GPU; the
logs are kept in gpurun_out/ (copied to profiles/ by the round's profiling script)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_check_under_torchrun(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    log = r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"dist_check_world{world}.log"), "w") as f:
            f.write(log)
    print(r.stdout[-3000:])
    assert r.returncode == 0, log[-3000:]
    assert f"DIST CHECK PASS (world {world}" in r.stdout
