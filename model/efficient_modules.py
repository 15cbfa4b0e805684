from constant_memory_waveglow_b200.efficient_modules import (  # noqa: F401
    AffineCouplingBlock, AffineCouplingFunc, Conv1x1Func, InvAffineCouplingFunc, InvConv1x1Func, InvertibleConv1x1)

__all__ = ['InvertibleConv1x1', 'AffineCouplingBlock']
