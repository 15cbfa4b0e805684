"""Direction dispatch and the flow-model contract: the surface of the reference's ``model/base.py``
(``Reversible`` :7-28, ``FlowBase`` :31-55) -- same names, methods and argument meaning.

A ``Reversible`` owns two computations, one the inverse of the other; ``reverse_mode=True`` swaps which of them ``forward()``
runs, so that a flow can be trained in the synthesis direction.  ``FlowBase`` adds the model contract
``forward(x, h) -> (z, logdet)``, ``reverse(z, h) -> (x, logdet)`` and ``infer(h, sigma) -> audio``.
"""
from typing import Callable, Optional, Tuple

import torch
from torch import Tensor, nn

FlowOut = Tuple[Tensor, Tensor]


class Reversible(nn.Module):
    _reverse_mode: bool

    def __init__(self, reverse_mode, **kwargs) -> None:
        super().__init__(**kwargs)
        self._reverse_mode = reverse_mode

    # the two directions a subclass provides
    def forward_computation(self, x: Tensor, *args, **kwargs) -> FlowOut:
        raise NotImplementedError(f"{type(self).__name__} defines no forward computation")

    def reverse_computation(self, z: Tensor, *args, **kwargs) -> FlowOut:
        raise NotImplementedError(f"{type(self).__name__} defines no reverse computation")

    def _direction(self, inverse: bool) -> Callable[..., FlowOut]:
        """The computation that realises the requested direction under this module's mode (XOR of the two switches)."""
        return self.reverse_computation if inverse != bool(self._reverse_mode) else self.forward_computation

    def forward(self, x: Tensor, *args, **kwargs) -> FlowOut:
        return self._direction(False)(x, *args, **kwargs)

    def reverse(self, z: Tensor, *args, **kwargs) -> FlowOut:
        return self._direction(True)(z, *args, **kwargs)


class FlowBase(Reversible):
    def __init__(self, condition_hop_length: int, reverse_mode=False) -> None:
        super().__init__(reverse_mode=reverse_mode)
        self._hop_length = condition_hop_length

    def forward_computation(self, x: Tensor, h: Tensor) -> FlowOut:
        raise NotImplementedError(f"{type(self).__name__} defines no forward computation")

    def reverse_computation(self, z: Tensor, h: Tensor) -> FlowOut:
        raise NotImplementedError(f"{type(self).__name__} defines no reverse computation")

    @torch.no_grad()
    def infer(self, h: Tensor, sigma: float = 1., z: Optional[Tensor] = None) -> Tensor:
        """Synthesis (``model/base.py:42-55``): z ~ N(0, sigma^2) with ``frames * hop`` samples per item, pushed through the
        direction that maps noise to audio.  ``z`` may be supplied (already scaled) so that parity tests feed the oracle and
        this path identical noise."""
        cond = h if h.dim() == 3 else h.unsqueeze(0)
        if z is None:
            n_items, frames = cond.shape[0], cond.shape[2]
            z = cond.new_empty((n_items, frames * self._hop_length)).normal_(std=sigma)
        audio, _ = self._direction(True)(z, cond)
        return audio.squeeze()
