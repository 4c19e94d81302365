"""Mirror of the reference's ``lib`` package for the hot path (UNet, Trainer, predict_linear_blend)."""
from . import AverageMeter, data_normalization, optim  # noqa: F401
from .UNet import UNet  # noqa: F401
