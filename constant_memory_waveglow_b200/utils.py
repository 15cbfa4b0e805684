"""Weight-norm helpers and the reflection factory (reference ``utils.py:5-16``).

``add_weight_norms`` fixes the parameter layout the autograd Functions see (``weight_g`` of shape
(out,1,1) + ``weight_v``); ``remove_weight_norms`` collapses them back to ``weight`` for inference.
The kernels never call the wrapped module's forward, so the weight-norm pre-hook costs nothing:
g*v/||v|| is recomputed on the device by ``cmwg_wn_pack`` once per weight version.
"""
import os

from torch import nn


def get_instance(module, config, *args):
    return getattr(module, config['type'])(*args, **config['args'])


def remove_weight_norms(m):
    if hasattr(m, 'weight_g'):
        nn.utils.remove_weight_norm(m)


def add_weight_norms(m):
    if hasattr(m, 'weight'):
        nn.utils.weight_norm(m)


def ensure_dir(path):
    if not os.path.exists(path):
        os.makedirs(path)
