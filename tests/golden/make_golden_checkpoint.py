"""Generate tests/golden/ref_map_golden.th with THE REFERENCE'S OWN GaussianMap.save()
(mapping/gaussian_map.py:491-507; build container only -- /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_checkpoint.py

The file is what a mission of the reference leaves on disk (`map_{idx}.th`); tests/test_checkpoint.py
loads it with the product's GaussianMap.load() and, where /root/reference is present, loads a file
written by the product's save() with the reference's load().  Alongside the checkpoint the script stores
what the reference's get_attr() returns for that map (ref_map_golden_attr.pt), so the product's
activations can be checked on the loaded state.
"""
import os
import sys
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from active_gs_b200 import synthetic as syn  # noqa: E402


def main():
    ops, mutils, gmap = mg.import_reference()
    gm = gmap.GaussianMap(mg.cfg_namespace(), "cpu")
    state = syn.make_room_scene(700, box=(3.0, 2.5, 2.0), seed=123, furniture=2)
    mg.load_state(gm, state)
    gm.is_init = True
    gm.save(HERE, index="golden")
    os.replace(os.path.join(HERE, "map_golden.th"), os.path.join(HERE, "ref_map_golden.th"))
    attr = [t.detach().clone() for t in gm.get_attr()]
    torch.save({"attr": attr, "normals": gm.get_normals.detach().clone()}, os.path.join(HERE, "ref_map_golden_attr.pt"))
    print("wrote", os.path.getsize(os.path.join(HERE, "ref_map_golden.th")), "bytes")


if __name__ == "__main__":
    main()
