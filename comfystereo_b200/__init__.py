"""comfystereo_b200 -- B200-native (sm_100a) implementation of ComfyStereo's depth-image-based
stereo generation hot path, packaged as the same ComfyUI custom node (reference __init__.py:11-55).

Layout:
  csrc/                     hand-written CUDA kernels + the C ABI (include/comfystereo_b200.h)
  _lib.py, engine.py        ctypes binding, workspaces, host/device entry points, frame sharding
  stereoimage_generation.py create_stereoimages / create_stereoimages_gpu (reference signatures)
  GenerateStereo.py         StereoImageNode (reference widget schema and outputs)
  synthetic.py              seeded synthetic frames shared by tests and bench
"""
from .GenerateStereo import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS, StereoImageNode  # noqa: F401

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS", "StereoImageNode"]
__version__ = "0.1.0"
