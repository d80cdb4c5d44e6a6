"""Renderer wrappers with the MONOCULAR tree's defaults (/root/reference/monocular/nnutils/nmr.py:56-290): the camera
looks from offset_z = 5.0 (:164; the multiframe tree uses 0.0, multiframe/nnutils/nmr.py:119 — mask and pix_to_face are the
same, zbuf is shifted).  Point monocular/nnutils/nmr.py at this module."""
from . import nmr


class NeuralRenderer(nmr.NeuralRenderer):
    def __init__(self, img_size=256, offset_z=5.0):
        super().__init__(img_size, offset_z)


OF_NeuralRenderer = nmr.OF_NeuralRenderer
