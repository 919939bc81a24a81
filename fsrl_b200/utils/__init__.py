"""Host-side utilities with the reference's export surface (fsrl/utils/__init__.py: the loggers and the PID
dual optimiser; ``BasicLogger`` is listed in the reference's ``__all__`` but defined nowhere)."""
from .logger import BaseLogger, DummyLogger, TensorboardLogger, WandbLogger
from .optim_util import LagrangianOptimizer

__all__ = ["BaseLogger", "TensorboardLogger", "DummyLogger", "WandbLogger", "LagrangianOptimizer"]
