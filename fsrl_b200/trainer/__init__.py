from .base_trainer import BaseTrainer
from .offpolicy import OffpolicyTrainer, offpolicy_trainer
from .onpolicy import OnpolicyTrainer, onpolicy_trainer

__all__ = ["BaseTrainer", "OnpolicyTrainer", "OffpolicyTrainer", "onpolicy_trainer", "offpolicy_trainer"]
