from .Contrast import MemoryMoCo
from .NCECriterion import NCESoftmaxLoss, fused_prob

__all__ = ["MemoryMoCo", "NCESoftmaxLoss", "fused_prob"]
