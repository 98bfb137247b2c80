"""Build libags_b200.so in-tree with nvcc for sm_100a (no torch headers involved)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
OUT = os.path.join(HERE, "libags_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SRC + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [
        os.path.join(HERE, "..", "include", "ags_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = ["nvcc"] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
