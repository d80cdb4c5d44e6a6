"""acfm_video_3d_reconstruction_b200 — B200-native (sm_100a) render-and-reproject hot path of ACFM.

Host-side mirror of the reference's nnutils API for that path (nmr.py, geom_utils.py,
loss_utils.py) over libacfm_b200.so (C ABI: include/acfm_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library was not built)
from . import camera, deform, functional, geom_utils, graphs, image_utils, loss_utils, monocular, multiframe, nmr, parallel, texture  # noqa: F401
from .nmr import NeuralRenderer, OF_NeuralRenderer  # noqa: F401

__version__ = "0.1.0"
