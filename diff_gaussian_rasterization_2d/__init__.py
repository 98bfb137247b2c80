"""Drop-in replacement for the un-vendored CUDA extension the reference imports at
/root/reference/utils/operations.py:22-25 -- same module name, same two symbols, backed by
libags_b200.so (sm_100a).  Put the repo root on PYTHONPATH and the reference runs unchanged."""
from active_gs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
