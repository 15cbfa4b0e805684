from constant_memory_waveglow_b200.condition import *  # noqa: F401,F403
from constant_memory_waveglow_b200.condition import MelSpec  # noqa: F401
