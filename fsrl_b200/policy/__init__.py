from .base_policy import ActorCritic, BasePolicy, DeviceBatch
from .lagrangian_base import LagrangianPolicy
from .ppo_lag import PPOLagrangian

__all__ = ["ActorCritic", "BasePolicy", "DeviceBatch", "LagrangianPolicy", "PPOLagrangian"]
