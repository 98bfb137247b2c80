"""GPU: the tuning variants of composite_bwd stay correct.  The variant is chosen by environment variables
that the library reads once per process (AGS_BWD_RED = reduction flavour, AGS_BWD_PX = pixels per lane,
AGS_BWD_TMA = TMA bulk staging of the depth-sorted records), so every variant runs the oracle parity tests
of tests/test_rasterizer_gpu.py in its own subprocess."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SELECT = "test_c1_forward_backward or test_room_cut_20k_160x120 or test_oversize_tile_merge_sort_path"


@pytest.mark.parametrize("env", [
    {"AGS_BWD_TMA": "1"},                       # cp.async.bulk + mbarrier staging (default reduction)
    {"AGS_BWD_RED": "0"}, {"AGS_BWD_RED": "1"}, {"AGS_BWD_RED": "2"}, {"AGS_BWD_RED": "3"},
    {"AGS_BWD_RED": "4"},                       # tensor-core (mma.sync TF32 hi/lo) reduction
    {"AGS_BWD_PX": "2", "AGS_BWD_RED": "2"}, {"AGS_BWD_PX": "4", "AGS_BWD_RED": "0"},
], ids=lambda e: ",".join(f"{k[8:]}={v}" for k, v in e.items()))
def test_backward_variant_parity(env):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    full = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_rasterizer_gpu.py"), "-m", "gpu",
                        "-q", "--tb=short", "-k", SELECT], cwd=ROOT, env=full, capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert "3 passed" in r.stdout
