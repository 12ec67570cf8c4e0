"""rnnlogic_b200 -- B200-native (sm_100a) reasoning-predictor hot path of RNNLogic.

Only what the hot path needs: csrc/ (CUDA kernels + C-ABI, include/rnnlogic_b200.h) and the
host-side mirror of the reference's Python interface for this path (KnowledgeGraph, Predictor,
PredictorPlus, TrainerPredictor, datasets)."""
from .graph import KnowledgeGraph  # noqa: F401
from .rules import CompiledRules, parse_rules  # noqa: F401

__all__ = ["KnowledgeGraph", "CompiledRules", "parse_rules"]
