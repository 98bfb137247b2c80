"""ctypes binding of libags_b200.so (C ABI in include/ags_b200.h).

The library is the product: there is NO CPU or PyTorch fallback.  If the shared object is missing
or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AGS_B200_LIB selects an alternative build of the same library (tuning experiments only)
LIB_PATH = os.environ.get("AGS_B200_LIB") or os.path.join(_HERE, "libags_b200.so")

AGS_NUM_STATS = 72
STAT_INSTANCES, STAT_OVERFLOW, STAT_VISIBLE, STAT_VIEW0 = 0, 1, 2, 8
PARAMS_ACTIVATED, PARAMS_RAW = 0, 1
ADAM_GROUPS = 5
CAM_ROW = 34

_f = C.c_void_p  # device pointers travel as void*


class RenderArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("param_mode", C.c_int32), ("require_importance", C.c_int32), ("front_only", C.c_int32),
        ("inst_cap", C.c_int32),
        ("scale_modifier", C.c_float), ("weight_thres", C.c_float), ("scale_factor", C.c_float),
        ("scale_max", C.c_float),
        ("means3D", _f), ("scales", _f), ("rotations", _f), ("opacities", _f), ("colors", _f),
        ("confidences", _f),
        ("viewmatrix", _f), ("projmatrix", _f), ("tanfov", _f), ("bg", _f), ("render_mask", _f),
        ("out_rgb", _f), ("out_normal", _f), ("out_depth", _f), ("out_opacity", _f),
        ("out_confidence", _f), ("importance", _f), ("count", _f), ("radii", _f), ("stats", _f),
        ("workspace", _f), ("workspace_bytes", C.c_size_t), ("stream", _f),
    ]


class RenderGradArgs(C.Structure):
    _fields_ = [
        ("d_rgb", _f), ("d_normal", _f), ("d_depth", _f), ("d_opacity", _f), ("d_confidence", _f),
        ("d_means3D", _f), ("d_scales", _f), ("d_rotations", _f), ("d_opacities", _f),
        ("d_colors", _f), ("d_means2D", _f), ("accumulate", C.c_int32), ("clear_records", C.c_int32),
    ]


class LossArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("B_total", C.c_int32),
        ("rgb", _f), ("normal", _f), ("depth", _f), ("opacity", _f),
        ("rgb_gt", _f), ("depth_gt", _f), ("tanfov", _f), ("vis_count", _f),
        ("normal_unit", _f), ("d2n", _f), ("d_rgb", _f), ("d_normal", _f), ("d_depth", _f),
        ("loss_terms", _f),
        ("w_depth", C.c_float), ("w_cons", C.c_float), ("w_tv", C.c_float),
        ("workspace", _f), ("workspace_bytes", C.c_size_t), ("stream", _f),
        ("frame_weight", _f), ("rgb_gt_frames_host", C.POINTER(C.c_void_p)),
        ("depth_gt_frames_host", C.POINTER(C.c_void_p)),
    ]


class AdamArgs(C.Structure):
    _fields_ = [
        ("num_groups", C.c_int32),
        ("param", _f * ADAM_GROUPS), ("grad", _f * ADAM_GROUPS), ("exp_avg", _f * ADAM_GROUPS),
        ("exp_avg_sq", _f * ADAM_GROUPS), ("numel", C.c_int64 * ADAM_GROUPS),
        ("lr", C.c_float * ADAM_GROUPS),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("step", C.c_int32), ("step_dev", _f), ("skip_flag", _f), ("stream", _f),
        ("zero_grad", C.c_int32),
    ]


MAX_PEERS = 8


class DistSync(C.Structure):
    _fields_ = [("peers", _f * 8), ("epoch", C.c_int32)]


SYNC_VIS, SYNC_TERMS, SYNC_GRADS, SYNC_PARAMS, SYNC_WORDS = 0, 1, 2, 3, 64


class DistAdamArgs(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("num_groups", C.c_int32), ("step", C.c_int32),
        ("grad_peers", _f * MAX_PEERS), ("param_peers", _f * MAX_PEERS), ("skip_peers", _f * MAX_PEERS),
        ("grad_multicast", _f), ("param_multicast", _f), ("exp_avg", _f), ("exp_avg_sq", _f),
        ("numel", C.c_int64 * ADAM_GROUPS), ("numel_padded", C.c_int64), ("lr", C.c_float * ADAM_GROUPS),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("stream", _f),
        ("sync", DistSync),
    ]


class DistVisArgs(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("opacity", _f), ("vis_local", _f), ("vis_peers", _f * MAX_PEERS), ("vis_multicast", _f),
        ("vis_count", _f), ("stream", _f), ("frame_weight", _f), ("sync", DistSync),
    ]


class DistTermsArgs(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("nterm", C.c_int32), ("nview", C.c_int32), ("terms", _f),
        ("stats", _f),
        ("gather_peers", _f * MAX_PEERS), ("gather_multicast", _f), ("stream", _f), ("sync", DistSync),
    ]


class SpawnArgs(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32),
        ("rgb", _f), ("depth", _f), ("depth_smooth", _f),
        ("c2w", C.c_float * 16), ("Kinv", C.c_float * 9),
        ("pred_rgb", _f), ("pred_depth", _f), ("pred_opacity", _f),
        ("error_thres", C.c_float), ("voxel_size", C.c_float), ("seed", C.c_uint32),
        ("n_old", C.c_int32), ("capacity", C.c_int32),
        ("means", _f), ("scales", _f), ("rotations", _f), ("opacities", _f), ("harmonics", _f),
        ("view_scores", _f), ("view_supports", _f), ("view_means", _f),
        ("counters", _f), ("select_out", _f),
        ("workspace", _f), ("workspace_bytes", C.c_size_t), ("stream", _f),
    ]


class PruneArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("T", C.c_int32), ("counts", _f), ("prune_mask", _f),
        ("src", _f * 8), ("dst", _f * 8), ("n_kept", _f),
        ("workspace", _f), ("workspace_bytes", C.c_size_t), ("stream", _f),
    ]


class UtilityArgs(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("M", C.c_int32),
        ("depth", _f), ("confidence", _f), ("valid", _f), ("voxel_centers", _f), ("unexplored", _f),
        ("w2c", _f), ("K", _f), ("depth_lo", C.c_float), ("depth_hi", C.c_float),
        ("explore", _f), ("exploit", _f), ("stream", _f),
    ]


class VoxelRoiArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("means", _f), ("rotations", _f), ("opacities", _f), ("confidences", _f),
        ("bbox_min", C.c_float * 3), ("voxel_size", C.c_float * 3), ("dim", C.c_int32 * 3),
        ("confidence_thres", C.c_float), ("opacity_thres", C.c_float), ("min_gaussian_per_voxel", C.c_int32),
        ("voxel_count", _f), ("voxel_normal", _f), ("update_mask", _f), ("stream", _f),
    ]


_lib = None

EXPORTS = ["ags_scratch_bytes", "ags_render_forward", "ags_render_backward", "ags_render_stage",
           "ags_loss_scratch_bytes", "ags_loss_forward_backward", "ags_postprocess", "ags_adam_step",
           "ags_dist_adam_step", "ags_smooth_depth", "ags_stage_cameras",
           "ags_spawn_scratch_bytes", "ags_spawn", "ags_view_stats_update", "ags_prune_scratch_bytes",
           "ags_prune_compact", "ags_view_utility", "ags_dist_vis_local", "ags_dist_vis_sum", "ags_dist_terms_put",
           "ags_voxel_roi", "ags_dist_wait",
           "ags_last_error", "ags_version", "ags_launch_count"]


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.ags_scratch_bytes.restype = C.c_size_t
    lib.ags_scratch_bytes.argtypes = [C.c_int32] * 5
    lib.ags_loss_scratch_bytes.restype = C.c_size_t
    lib.ags_loss_scratch_bytes.argtypes = [C.c_int32] * 3
    lib.ags_render_forward.argtypes = [C.POINTER(RenderArgs)]
    lib.ags_render_backward.argtypes = [C.POINTER(RenderArgs), C.POINTER(RenderGradArgs)]
    lib.ags_render_stage.argtypes = [C.POINTER(RenderArgs), C.POINTER(RenderGradArgs), C.c_int]
    lib.ags_render_stage.restype = C.c_int
    lib.ags_loss_forward_backward.argtypes = [C.POINTER(LossArgs)]
    lib.ags_adam_step.argtypes = [C.POINTER(AdamArgs)]
    lib.ags_smooth_depth.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                     C.c_float, C.c_void_p, C.c_void_p]
    lib.ags_smooth_depth.restype = C.c_int
    lib.ags_stage_cameras.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ags_stage_cameras.restype = C.c_int
    lib.ags_spawn_scratch_bytes.restype = C.c_size_t
    lib.ags_spawn_scratch_bytes.argtypes = [C.c_int32] * 2
    lib.ags_spawn.argtypes = [C.POINTER(SpawnArgs)]
    lib.ags_spawn.restype = C.c_int
    lib.ags_view_stats_update.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    lib.ags_view_stats_update.restype = C.c_int
    lib.ags_prune_scratch_bytes.restype = C.c_size_t
    lib.ags_prune_scratch_bytes.argtypes = [C.c_int32]
    lib.ags_prune_compact.argtypes = [C.POINTER(PruneArgs)]
    lib.ags_prune_compact.restype = C.c_int
    lib.ags_view_utility.argtypes = [C.POINTER(UtilityArgs)]
    lib.ags_view_utility.restype = C.c_int
    for name, typ in [("ags_dist_vis_local", DistVisArgs), ("ags_dist_vis_sum", DistVisArgs),
                      ("ags_dist_terms_put", DistTermsArgs)]:
        getattr(lib, name).argtypes = [C.POINTER(typ)]
        getattr(lib, name).restype = C.c_int
    lib.ags_voxel_roi.argtypes = [C.POINTER(VoxelRoiArgs)]
    lib.ags_voxel_roi.restype = C.c_int
    lib.ags_dist_wait.argtypes = [C.POINTER(DistSync), C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ags_dist_wait.restype = C.c_int
    lib.ags_dist_adam_step.argtypes = [C.POINTER(DistAdamArgs)]
    lib.ags_dist_adam_step.restype = C.c_int
    lib.ags_postprocess.argtypes = [C.c_int32] * 3 + [C.c_void_p] * 7
    lib.ags_postprocess.restype = C.c_int
    lib.ags_last_error.restype = C.c_char_p
    lib.ags_launch_count.restype = C.c_ulonglong
    lib.ags_launch_count.argtypes = []
    for name in ["ags_render_forward", "ags_render_backward", "ags_loss_forward_backward",
                 "ags_adam_step", "ags_version"]:
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ags_last_error().decode()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libags_b200 operates on CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return t.data_ptr()


def current_stream(device):
    return torch.cuda.current_stream(device).cuda_stream
