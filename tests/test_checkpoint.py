"""Row f4: the `.th` map checkpoint (mapping/gaussian_map.py:491-527) across the two implementations.

CPU (runs everywhere): the product's GaussianMap.load() reads tests/golden/ref_map_golden.th -- a file
WRITTEN BY THE REFERENCE'S OWN save() (tests/golden/make_golden_checkpoint.py) -- and reproduces the
reference's get_attr() on it; save() writes exactly N rows per tensor (no capacity-buffer storage).
Where /root/reference exists (the build container) the reference's own load() reads a file written by the
product's save().  GPU: the loaded map renders through the product's GaussianRenderer.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
KEYS = {"means", "scales", "harmonics", "opacities", "rotations", "view_scores", "view_supports",
        "view_means", "near", "far", "use_view_direction", "background_color", "scale_factor"}


def _product_map(device="cpu"):
    from active_gs_b200.gaussian_map import GaussianMap
    from active_gs_b200.config import default_gaussian_map_config
    return GaussianMap(default_gaussian_map_config(), device)


def test_product_loads_reference_written_checkpoint():
    ref_attr = torch.load(os.path.join(GOLD, "ref_map_golden_attr.pt"), weights_only=False)
    gm = _product_map()
    gm.load(os.path.join(GOLD, "ref_map_golden.th"))
    assert gm.is_init and gm._means.shape == (700, 3) and gm._harmonics.shape == (700, 1, 3)
    assert gm.scene_near == 0.001 and gm.scene_far == 10.0 and gm.scale_factor == 0.01
    for got, want in zip(gm.get_attr(), ref_attr["attr"]):
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(gm.get_normals, ref_attr["normals"], rtol=1e-6, atol=1e-7)


def test_product_save_writes_reference_keys_and_only_n_rows(tmp_path):
    gm = _product_map()
    gm.load(os.path.join(GOLD, "ref_map_golden.th"))
    # views of larger buffers, as after spawn/prune (capacity buffers) or fused multi-GPU training
    big = torch.zeros(100_000, 3)
    big[:700] = gm._means
    gm._means = big[:700]
    gm.save(str(tmp_path), index=3)
    path = tmp_path / "map_3.th"
    d = torch.load(path, weights_only=False)
    assert set(d.keys()) == KEYS
    ref = torch.load(os.path.join(GOLD, "ref_map_golden.th"), weights_only=False)
    for k in KEYS:
        a, b = d[k], ref[k]
        if torch.is_tensor(b):
            assert torch.equal(torch.as_tensor(a).cpu(), b.cpu()), k
        else:
            assert a == b, k
    # 76 B per Gaussian + pickle overhead, not the 1.2 MB of the capacity buffer
    assert os.path.getsize(path) < 1.15 * os.path.getsize(os.path.join(GOLD, "ref_map_golden.th")) + 4096


@pytest.mark.skipif(not os.path.isdir("/root/reference/mapping"), reason="the reference tree exists in the build container only")
def test_reference_loads_product_written_checkpoint(tmp_path):
    sys.path.insert(0, GOLD)
    import make_golden as mg
    ops, mutils, gmap = mg.import_reference()
    gm = _product_map()
    gm.load(os.path.join(GOLD, "ref_map_golden.th"))
    gm._opacities = gm._opacities + 0.25                      # not just a byte copy of the fixture
    gm.save(str(tmp_path), index="p")
    ref = gmap.GaussianMap(mg.cfg_namespace(), "cpu")
    ref.load(str(tmp_path / "map_p.th"))
    for got, want in zip(ref.get_attr(), gm.get_attr()):
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)
    assert ref.is_init


@pytest.mark.gpu
def test_reference_checkpoint_renders_on_the_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from active_gs_b200 import operations as O, synthetic as syn
    dev = torch.device("cuda:0")
    gm = _product_map(dev)
    gm.load(os.path.join(GOLD, "ref_map_golden.th"))
    ext, K = syn.make_cameras(2, box=(3.0, 2.5, 2.0), H=48, W=64, hfov=70.0, seed=5)
    with torch.no_grad():
        out = O.GaussianRenderer(ext.to(dev), K.to(dev), gm.get_attr(), gm.background_color,
                                 (gm.scene_near, gm.scene_far), (48, 64), dev).render_view_all()
    assert out[0].shape == (2, 3, 48, 64) and torch.isfinite(out[0]).all() and float(out[3].max()) > 0.5
