"""ni_slam_b200 -- B200-native (sm_100a) implementation of NI-SLAM's tracking / loop-closure hot path.

Only what the path needs: csrc/ (CUDA kernels + C ABI), host/ (C++ shim with the reference's class signatures),
api.py (the same surface for Python) and build.py (in-tree nvcc build).  There is no CPU fallback.
"""
from .api import (CameraModel, CFConfig, CorrelationFlow, KeyframeSelectionConfig, TRACK_RESULT_DTYPE, Frame, LoopClosure, LoopClosureConfig, LoopClosureResult, LoopResultC, MapStitcher, NisError,
                  LIB_PATH, SYMBOLS, load_library, loop_reduce, DB_FULL, DB_SPECTRA, DB_IMAGE, SCAN_RECORD_DTYPE)

__all__ = ["CameraModel", "KeyframeSelectionConfig", "TRACK_RESULT_DTYPE", "CFConfig", "CorrelationFlow", "Frame", "LoopClosure", "LoopClosureConfig", "LoopClosureResult", "LoopResultC",
           "MapStitcher", "NisError", "LIB_PATH", "SYMBOLS", "load_library", "loop_reduce", "DB_FULL", "DB_SPECTRA", "DB_IMAGE", "SCAN_RECORD_DTYPE"]
