from .base_agent import BaseAgent, OffpolicyAgent, OnpolicyAgent
from .ppo_lag_agent import PPOLagAgent

__all__ = ["BaseAgent", "OffpolicyAgent", "OnpolicyAgent", "PPOLagAgent"]
