"""active_gs_b200 -- B200-native (sm_100a) rasterize-and-optimise hot path of ActiveGS.

Sub-modules are imported lazily: `synthetic` is pure CPU, everything that touches the CUDA
library goes through `active_gs_b200.lib` which raises loudly when libags_b200.so is missing."""
__version__ = "0.1.0"
