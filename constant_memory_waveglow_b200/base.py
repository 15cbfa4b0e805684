"""Direction-dispatching base classes, mirroring the reference's ``model/base.py`` surface
(``Reversible`` :7-28, ``FlowBase`` :31-55): same names, methods and argument meaning."""
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor


class Reversible(nn.Module):
    """A module with a forward and an inverse computation; ``reverse_mode=True`` swaps which one
    ``forward()`` runs (reference ``model/base.py:20-28``)."""
    _reverse_mode: bool

    def __init__(self, reverse_mode, **kwargs) -> None:
        super().__init__(**kwargs)
        self._reverse_mode = reverse_mode

    def forward_computation(self, x: Tensor, *args, **kwargs) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def reverse_computation(self, z: Tensor, *args, **kwargs) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def forward(self, x: Tensor, *args, **kwargs) -> Tuple[Tensor, Tensor]:
        fn = self.reverse_computation if self._reverse_mode else self.forward_computation
        return fn(x, *args, **kwargs)

    def reverse(self, z: Tensor, *args, **kwargs) -> Tuple[Tensor, Tensor]:
        fn = self.forward_computation if self._reverse_mode else self.reverse_computation
        return fn(z, *args, **kwargs)


class FlowBase(Reversible):
    """Flow model contract (reference ``model/base.py:31-55``): ``forward(x, h) -> (z, logdet)``,
    ``reverse(z, h) -> (x, logdet)``, ``infer(h, sigma) -> audio``."""

    def __init__(self, condition_hop_length: int, reverse_mode=False) -> None:
        super().__init__(reverse_mode=reverse_mode)
        self._hop_length = condition_hop_length

    def forward_computation(self, x: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def reverse_computation(self, z: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    @torch.no_grad()
    def infer(self, h: Tensor, sigma: float = 1., z: Tensor = None) -> Tensor:
        """Sample z ~ N(0, sigma^2) of length frames*hop and run the synthesis direction
        (reference ``model/base.py:42-55``).  ``z`` may be supplied (already scaled) so that parity
        tests feed the oracle and this path identical noise."""
        if h.dim() == 2:
            h = h.unsqueeze(0)
        batch, _, steps = h.shape
        if z is None:
            z = h.new_empty((batch, steps * self._hop_length)).normal_(std=sigma)
        if self._reverse_mode:
            x, _ = self.forward_computation(z, h)
        else:
            x, _ = self.reverse_computation(z, h)
        return x.squeeze()
