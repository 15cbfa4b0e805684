from constant_memory_waveglow_b200.trainer import (Callback, DeviceStatsMonitor,  # noqa: F401
                                                   LearningRateMonitor, ModelSummary)
