"""Host-side mirror of the utility evaluation of the reference's planners (SURVEY.md section 8 row
f1): planning/confidence.py:6-101 (class Confidence) and planning/exploration.py:6-86 (class
Exploration), `cal_utility` only -- candidate sampling, path planning and the GUI queues of
planning/plan_base.py are outside the hot path and are not rebuilt.

The reference renders the ~100 candidate views one by one and evaluates each with ~40 small ATen
launches plus a nonzero() sync; here all candidates are rendered by ONE rasterizer launch per chunk
and evaluated by ONE kernel (ags_view_utility): visible-and-unexplored voxel fraction and mean
distance-weighted uncertainty per view, no host round trip until the final (V,) result.
"""
import time

import numpy as np
import torch

from . import ops
from . import operations as O
from .rasterizer import views_per_chunk


class _UtilityBase:
    CHUNK = 128            # candidate views per rasterizer launch (upper bound; see _chunk)
    WORKSPACE_BUDGET = 2 << 30

    def _chunk(self, gaussian_map, h, w):
        """views per launch from a workspace byte budget: 128 views over a 1 M-surfel map would ask
        for > 10 GB of per-(view, Gaussian) scratch before any visibility is known"""
        return views_per_chunk(gaussian_map.get_means.shape[0], h, w, self.WORKSPACE_BUDGET,
                               device_index=self.device.index or 0, cap=self.CHUNK)

    def __init__(self, cfg, device):
        self.device = torch.device(device)
        self.render_ratio = cfg.render_ratio

    def _render(self, gaussian_map, extrinsics, intrinsics, h, w):
        depth, conf = [], []
        attrs = gaussian_map.get_attr()
        chunk = self._chunk(gaussian_map, h, w)
        for c0 in range(0, extrinsics.shape[0], chunk):
            out = O.GaussianRenderer(extrinsics[c0:c0 + chunk], intrinsics[c0:c0 + chunk], attrs,
                                     gaussian_map.background_color, (gaussian_map.scene_near, gaussian_map.scene_far),
                                     (h, w), self.device).render_view_all()
            depth.append(out[1][:, 0])
            conf.append(out[5][:, 0])
        return torch.cat(depth).contiguous(), torch.cat(conf).contiguous()

    def _valid_masks(self, simulator, extrinsics, h, w):
        """planning/confidence.py:52-66: the simulator's valid-surface mask per candidate, nearest-
        neighbour resized to the render resolution (only for datasets with missing surfaces)."""
        if not getattr(simulator, "has_missing_surface", False):
            return None, 0.0
        import cv2
        t0 = time.time()
        masks = []
        for i in range(extrinsics.shape[0]):
            m = simulator.simulate(extrinsics[i].cpu(), valid_mask_only=True)
            masks.append(cv2.resize(m.astype(np.uint8), (int(h), int(w)), interpolation=cv2.INTER_NEAREST))
        return torch.tensor(np.stack(masks)).to(self.device).contiguous(), time.time() - t0

    @torch.no_grad()
    def _utilities(self, gaussian_map, voxel_map, candidates, simulator):
        t0 = time.time()
        h, w = (int(v) for v in np.round(self.render_ratio * np.asarray(simulator.resolution)).astype(int))
        extrinsics = candidates.to(self.device).float()
        intrinsics = simulator.intrinsic.to(self.device).float()[None].expand(extrinsics.shape[0], 3, 3)
        depth, conf = self._render(gaussian_map, extrinsics, intrinsics, h, w)
        valid, t_sim = self._valid_masks(simulator, extrinsics, h, w)
        explore, exploit = ops.view_utility(
            depth, conf, voxel_map.voxel_centers.to(self.device), voxel_map.unexplored_mask.to(self.device),
            torch.linalg.inv(extrinsics.cpu()).to(self.device), intrinsics, simulator.depth_range, valid=valid)
        explore, exploit = explore.cpu(), exploit.cpu()              # the only synchronisation
        return explore, exploit, time.time() - t0 - t_sim


class Confidence(_UtilityBase):
    """planning/confidence.py:6-101."""

    def __init__(self, cfg, device):
        super().__init__(cfg, device)
        self.explore_weight = cfg.explore_weight

    def cal_utility(self, gaussian_map, voxel_map, candidates, simulator):
        explore, exploit, t = self._utilities(gaussian_map, voxel_map, candidates, simulator)
        return self.explore_weight * explore + exploit, t


class Exploration(_UtilityBase):
    """planning/exploration.py:6-86."""

    def cal_utility(self, gaussian_map, voxel_map, candidates, simulator):
        explore, _, t = self._utilities(gaussian_map, voxel_map, candidates, simulator)
        return explore, t


def low_confidence_voxels(voxel_map, gaussian_map, confidence_thres=0.3):
    """The Gaussian half of VoxelMap.update_utility (mapping/voxel_map.py:70-113) in one scatter kernel
    (the reference: 4 activation passes, 6 boolean-index compactions, 2 scatter_adds): returns
    (voxel_normal (M,3), update_mask (M,) bool) for `self.voxel_normal` / `raw_roi_mask += update_mask`.
    `voxel_map` needs bbox (2,3), size (3,), dim (3,), min_gaussian_per_voxel."""
    gm = gaussian_map
    _, normal, mask = ops.voxel_roi(gm.get_means.detach(), gm._rotations.detach(), gm._opacities.detach(),
                                    gm.get_confidences.detach(), voxel_map.bbox[0].tolist(), voxel_map.size.tolist(),
                                    [int(d) for d in voxel_map.dim], confidence_thres=confidence_thres,
                                    min_gaussian_per_voxel=voxel_map.min_gaussian_per_voxel)
    return normal, mask


def dilate6(mask3d):
    """6-neighbourhood binary dilation of a (X,Y,Z) bool tensor on its own device -- what
    scipy.ndimage.binary_dilation(mask, generate_binary_structure(3, 1)) returns
    (mapping/voxel_map.py:21,223-229,290-304), without the numpy round trip of the reference."""
    m = mask3d.bool()
    out = m.clone()
    out[1:, :, :] |= m[:-1, :, :]; out[:-1, :, :] |= m[1:, :, :]
    out[:, 1:, :] |= m[:, :-1, :]; out[:, :-1, :] |= m[:, 1:, :]
    out[:, :, 1:] |= m[:, :, :-1]; out[:, :, :-1] |= m[:, :, 1:]
    return out


def update_utility(voxel_map, gaussian_map, use_confidence, confidence_thres=0.3):
    """VoxelMap.update_utility (mapping/voxel_map.py:62-116) for a reference-shaped voxel map object (attributes
    dim, bbox, size, min_gaussian_per_voxel, frontier_mask, free_mask): region-of-interest voxels = (frontier or
    holding more than min_gaussian_per_voxel opaque low-confidence Gaussians) and next to free space.  The
    Gaussian half is one scatter kernel (low_confidence_voxels), the dilation stays on the device.  Sets
    voxel_map.voxel_normal / voxel_map.roi_mask like the reference and returns the mask."""
    dim = [int(d) for d in voxel_map.dim]
    M = dim[0] * dim[1] * dim[2]
    dev = gaussian_map.device
    voxel_map.voxel_normal = torch.zeros((M, 3), device=dev)
    raw = voxel_map.frontier_mask.to(dev).bool().clone()
    if use_confidence:
        normal, upd = low_confidence_voxels(voxel_map, gaussian_map, confidence_thres)
        voxel_map.voxel_normal = normal
        raw |= upd
    free = voxel_map.free_mask.to(dev).view(*dim)
    voxel_map.roi_mask = raw & dilate6(free).view(-1)
    return voxel_map.roi_mask
