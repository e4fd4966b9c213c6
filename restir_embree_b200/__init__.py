"""restir_embree_b200 — B200-native ReSTIR DI hot path (drop-in for Tonz24/restir-embree's
produceRestir + Embree ray queries). The compute lives in csrc/ (hand-written CUDA for sm_100a
behind the C ABI of include/restir_b200.h); this package is the thin host-side mirror."""
from . import abi, camera, scenes  # noqa: F401
from .camera import Camera  # noqa: F401

__all__ = ["abi", "camera", "scenes", "Camera"]
