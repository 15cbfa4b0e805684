from constant_memory_waveglow_b200.trainer import DDPPlugin  # noqa: F401
