"""Counterpart of the one numerical routine of ``morgana.viz`` that sits on the models' ``predict`` path: MLPG."""
from morgana_b200.viz import synthesis   # noqa: F401
