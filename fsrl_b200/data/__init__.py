from .batch import Batch, to_numpy, to_torch_as
from .buffer import DeviceVectorReplayBuffer, VectorReplayBuffer
from .fast_collector import FastCollector

__all__ = ["Batch", "to_numpy", "to_torch_as", "DeviceVectorReplayBuffer", "VectorReplayBuffer",
           "FastCollector"]
