from constant_memory_waveglow_b200.waveflow import WN2D, NonCausalLayer2D, WaveFlow  # noqa: F401  (reference model/waveflow.py)
