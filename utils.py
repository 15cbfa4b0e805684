"""Top-level ``utils`` module the reference scripts import (reference ``utils.py``)."""
from constant_memory_waveglow_b200.utils import (add_weight_norms, ensure_dir, get_instance,  # noqa: F401
                                                 remove_weight_norms)
