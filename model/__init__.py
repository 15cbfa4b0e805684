"""Drop-in ``model`` package: the import names the reference's train.py / inference.py /
tests/test_fwd_bwd.py use (reference ``model/__init__.py:1-7``), re-exported from
constant_memory_waveglow_b200."""
from constant_memory_waveglow_b200.base import FlowBase, Reversible
from constant_memory_waveglow_b200.waveflow import WaveFlow
from constant_memory_waveglow_b200.waveglow import WaveGlow
from constant_memory_waveglow_b200.wsrglow import WSRGlow
from constant_memory_waveglow_b200.mr_waveglow import MRWaveGlow
from constant_memory_waveglow_b200.melglow import MelGlow
from constant_memory_waveglow_b200.trainer import LightModel
from . import condition  # noqa: F401  (reference: `from model import LightModel, condition`, inference.py:10)

