"""Import-name shim: ``pytorch_lightning`` is not installed in this image (nor in its wheelhouse) and the
reference's ``train.py:8-10`` / ``model/lightning.py:5`` import it.  The names they use resolve to the Lightning-free
harness in ``constant_memory_waveglow_b200.trainer``.  With the real package installed, delete this directory."""
from constant_memory_waveglow_b200.trainer import (Callback, LightningModule, Trainer,  # noqa: F401
                                                   seed_everything)
from . import callbacks, plugins  # noqa: F401

__version__ = "0+cmwg_b200"
