"""Reference ``model/lightning.py``: ``LightModel`` on the Lightning-free harness."""
from constant_memory_waveglow_b200.trainer import LightModel  # noqa: F401
