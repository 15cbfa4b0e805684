"""Drop-in ``datasets`` package (the reference's git submodule): ``model/lightning.py:11,47`` builds the training set
by reflection from here."""
from constant_memory_waveglow_b200.datasets import RandomWAVDataset  # noqa: F401
