"""CPU: the C-ABI library loads, exports every symbol include/ags_b200.h declares, the ctypes
mirrors in active_gs_b200/lib.py have the same layout as the C structs (checked with gcc), and the
argument validation / error plumbing works without a GPU (no compute calls)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "ags_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from active_gs_b200 import lib as L
    return L


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ags_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    h = lib.load()
    names = declared_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(h, n), f"{n} declared in ags_b200.h but not exported"
    assert set(lib.EXPORTS) == set(names)
    assert h.ags_version() == 100


def test_struct_layouts_match_c(lib, tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ags_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(AgsRenderArgs),sizeof(AgsRenderGradArgs),sizeof(AgsLossArgs),sizeof(AgsAdamArgs),'
                    'offsetof(AgsRenderArgs,workspace_bytes),offsetof(AgsLossArgs,workspace),offsetof(AgsAdamArgs,skip_flag),'
                    'sizeof(AgsSpawnArgs),sizeof(AgsPruneArgs),sizeof(AgsUtilityArgs),offsetof(AgsSpawnArgs,counters),'
                    'offsetof(AgsPruneArgs,n_kept),offsetof(AgsUtilityArgs,explore));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(lib.RenderArgs), C.sizeof(lib.RenderGradArgs), C.sizeof(lib.LossArgs), C.sizeof(lib.AdamArgs),
            lib.RenderArgs.workspace_bytes.offset, lib.LossArgs.workspace.offset, lib.AdamArgs.skip_flag.offset,
            C.sizeof(lib.SpawnArgs), C.sizeof(lib.PruneArgs), C.sizeof(lib.UtilityArgs), lib.SpawnArgs.counters.offset,
            lib.PruneArgs.n_kept.offset, lib.UtilityArgs.explore.offset]
    assert got == want


def test_scratch_size_query_and_argument_errors(lib):
    h = lib.load()
    small = h.ags_scratch_bytes(1000, 1, 64, 64, 1 << 16)
    big = h.ags_scratch_bytes(200000, 8, 480, 640, 8 << 20)
    assert 0 < small < big
    assert h.ags_scratch_bytes(-1, 1, 64, 64, 0) == 0
    # per Gaussian-view: 4 records of 16 B + rect 8 B + visible-list entry 4 B + gradient record 64 B
    d = h.ags_scratch_bytes(2000, 1, 64, 64, 1 << 16) - small
    assert abs(d - 1000 * (64 + 8 + 4 + 64)) <= 7 * 256
    a = lib.RenderArgs()
    a.N, a.B, a.H, a.W = 10, 0, 64, 64
    rc = h.ags_render_forward(C.byref(a))
    assert rc < 0 and b"bad sizes" in h.ags_last_error()
    assert h.ags_render_forward(None) < 0
    ad = lib.AdamArgs()
    ad.num_groups = 9
    assert h.ags_adam_step(C.byref(ad)) < 0 and b"num_groups" in h.ags_last_error()
    lo = lib.LossArgs()
    assert h.ags_loss_forward_backward(C.byref(lo)) < 0


def test_product_refuses_cpu_tensors(lib):
    """There is no CPU fallback: CPU tensors raise instead of silently computing elsewhere."""
    from active_gs_b200.rasterizer import RenderBatch
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA"):
        RenderBatch(z(4, 3), z(4, 3), z(4, 4), z(4), z(4, 3), z(4), torch.eye(4)[None], torch.eye(4)[None],
                    z(1, 2), z(3), 16, 16)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure and baseline/ is bench-only: no module of the product packages may import them."""
    for pkg in ["active_gs_b200", "diff_gaussian_rasterization_2d"]:
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{pkg}/{f} imports oracle"
                    assert not re.search(r"^\s*(from|import)\s+baseline\b", src, flags=re.M), f"{pkg}/{f} imports baseline"
