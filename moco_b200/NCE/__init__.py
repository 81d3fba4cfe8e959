from .Contrast import MemoryMoCo
from .NCECriterion import NCESoftmaxLoss, fused_prob
from .ShardedContrast import ShardedMemoryMoCo

__all__ = ["MemoryMoCo", "NCESoftmaxLoss", "ShardedMemoryMoCo", "fused_prob"]
