"""The hot-path constants of the reference's hydra config as plain attribute objects
(/root/reference/config/mapper/incremental.yaml:12-32).  hydra/omegaconf are not needed."""
from types import SimpleNamespace as NS


def default_gaussian_map_config(**over):
    cfg = NS(bound=[0.001, 10.0], background=[0.0, 0.0, 0.0, 0.0], sparse_ratio=0.1, error_thres=0.25,
             scale_factor=0.01, optimization_steps=10, prune_interval=5, use_view_distribution=True,
             sampler=NS(sampler_type="weighted", batch_size=8, active_size=3),
             optimizer=NS(mean_lr=0.0005, rotation_lr=0.0005, opacity_lr=0.01, scale_lr=0.01,
                          harmonic_lr=0.0001))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
