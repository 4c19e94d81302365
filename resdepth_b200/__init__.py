"""resdepth_b200 -- B200-native (sm_100a) implementation of the ResDepth hot path.

``resdepth_b200.lib`` mirrors the reference's ``lib`` package for the path: ``lib.UNet.UNet``,
``lib.Trainer.Trainer``, ``lib.evaluation.predict_linear_blend``; everything runs through the C-ABI library
``resdepth_b200/_lib/libresdepth_b200.so`` (``include/resdepth_b200.h``).
"""
__version__ = '0.1.0'

from . import _native  # noqa: F401
