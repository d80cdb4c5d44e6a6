"""Renderer wrappers with the MULTIFRAME tree's defaults (/root/reference/multiframe/nnutils/nmr.py:54-238: offset_z = 0.0,
:119).  Point multiframe/nnutils/nmr.py at this module."""
from .nmr import NeuralRenderer, OF_NeuralRenderer  # noqa: F401
