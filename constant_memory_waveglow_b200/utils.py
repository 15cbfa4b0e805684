"""Weight-norm helpers and the reflection factory: the names and behaviour of the reference's ``utils.py:5-16``.

``add_weight_norms`` / ``remove_weight_norms`` are ``module.apply`` visitors.  They fix the parameter layout the autograd
Functions see (``weight_g`` of shape (out, 1, 1) + ``weight_v``) and collapse it back to ``weight`` for inference.  The kernels
never call a wrapped module's forward, so the weight-norm pre-hook costs nothing at run time: g * v / ||v|| is recomputed on
the device by ``cmwg_wn_pack`` once per weight version.
"""
from pathlib import Path
from typing import Any, Mapping

from torch.nn.utils import remove_weight_norm, weight_norm

__all__ = ["get_instance", "add_weight_norms", "remove_weight_norms", "ensure_dir"]


def get_instance(module: Any, config: Mapping[str, Any], *args):
    """Build ``module.<config['type']>(*args, **config['args'])`` -- how ``LightModel`` turns the JSON blocks ``arch``,
    ``conditioner``, ``loss``, ``optimizer`` and ``dataset`` into objects (``model/lightning.py:33-35,42-49``)."""
    factory = getattr(module, config["type"])
    return factory(*args, **config["args"])


def add_weight_norms(m) -> None:
    """Give every visited module that owns a ``weight`` the (``weight_g``, ``weight_v``) parametrisation over dim 0."""
    if not hasattr(m, "weight"):
        return
    weight_norm(m, name="weight", dim=0)


def remove_weight_norms(m) -> None:
    """Fold (``weight_g``, ``weight_v``) back into a plain ``weight`` where a visited module has them."""
    if not hasattr(m, "weight_g"):
        return
    remove_weight_norm(m, name="weight")


def ensure_dir(path) -> None:
    Path(path).mkdir(parents=True, exist_ok=True)
