from constant_memory_waveglow_b200.melglow import WN_LVC, MelGlow, NonCausalLayerLVC, Predictor  # noqa: F401  (reference model/melglow.py)
