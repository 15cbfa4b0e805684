from constant_memory_waveglow_b200.wsrglow import AngleEmbedding, WSRGlow  # noqa: F401
