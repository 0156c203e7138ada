"""redtail_b200 -- B200-native stereo-depth inference behind the redtail stereoDNN plugin API.

Python here is a thin ctypes binding over the two in-tree shared libraries built from redtail_b200/csrc:

  lib/libredtail_b200.so        C-ABI over the hand-written sm_100a kernels   (include/redtail_b200.h)
  lib/libnvstereo_inference.so  nvinfer1-compatible engine + plugins + nets   (include/redtail_b200_engine.h)

PyTorch is used only for device memory and streams.  There is no CPU or eager fallback: importing works without
a GPU (so symbols can be inspected), every compute call fails loudly without one or without the built libraries.
"""
from ._lib import kernels_lib, engine_lib, lib_paths, LibraryMissing  # noqa: F401
from . import ops  # noqa: F401
from .engine import StereoEngine, CaffeNet  # noqa: F401

__all__ = ["ops", "StereoEngine", "CaffeNet", "kernels_lib", "engine_lib", "lib_paths", "LibraryMissing"]
