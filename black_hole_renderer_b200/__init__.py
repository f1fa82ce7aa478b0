"""B200-native (sm_100a) implementation of hwuu/black-hole-renderer's per-pixel null-geodesic
render path.  `Renderer` mirrors the reference's `TaichiRenderer`; the kernels live in libbhr.so
(csrc/, C-ABI in include/bhr.h)."""
from .camera import build_camera  # noqa: F401
from .renderer import Renderer, TaichiRenderer, compute_edge_alpha  # noqa: F401

__all__ = ["Renderer", "TaichiRenderer", "build_camera", "compute_edge_alpha"]
