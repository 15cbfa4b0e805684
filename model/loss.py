from constant_memory_waveglow_b200.loss import WaveGlowLoss  # noqa: F401
