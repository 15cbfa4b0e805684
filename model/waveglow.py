from constant_memory_waveglow_b200.waveglow import WN, NonCausalLayer, WaveGlow, fused_gate  # noqa: F401
