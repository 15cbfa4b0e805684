from constant_memory_waveglow_b200.base import FlowBase, Reversible  # noqa: F401
