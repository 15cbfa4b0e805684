from constant_memory_waveglow_b200.mr_waveglow import MRWaveGlow  # noqa: F401  (reference model/mr_waveglow.py)
