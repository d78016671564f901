"""WaveNet building blocks with the reference's constructor surface
(movenet/modules.py:15-142 of cosmicBboy/movenet).

In this package the blocks are parameter holders: they own ``nn.Conv1d``
tensors with exactly the reference's names, shapes and default initialisation,
so ``state_dict()`` round-trips with reference checkpoints in both directions.
The arithmetic does not run module by module -- ``WaveNet.forward`` hands the
whole stack to the CUDA library in one call -- so calling one of these blocks
on its own raises instead of silently running a PyTorch implementation.
"""
from typing import List

import torch.nn as nn


class _HotPathBlock(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} is evaluated inside movenet_b200.WaveNet's fused CUDA path; "
            "it has no standalone (PyTorch/CPU) forward")


class CausalConv1d(_HotPathBlock):
    """``Conv1d(k=2, padding=1, bias=False)`` minus the last column: movenet/modules.py:15-30."""

    def __init__(self, input_channels, out_channels, kernel_size=2, bias=False):
        super().__init__()
        if kernel_size != 2 or bias:
            raise ValueError("the hot path implements kernel_size=2, bias=False (the only use in wavenet.py:119)")
        self.kernel_size = kernel_size
        self.conv = nn.Conv1d(input_channels, out_channels, kernel_size=kernel_size, stride=1,
                              padding=kernel_size - 1, bias=bias)


class DilatedCausalConv1d(_HotPathBlock):
    """``Conv1d(k=2, dilation=d, bias=False)``: movenet/modules.py:33-46."""

    def __init__(self, channels, dilation=1, kernel_size=2, bias=False):
        super().__init__()
        if kernel_size != 2 or bias:
            raise ValueError("the hot path implements kernel_size=2, bias=False")
        self.conv = nn.Conv1d(channels, channels, kernel_size=kernel_size, stride=1, dilation=dilation, bias=bias)


class GatedResidualConv1d(_HotPathBlock):
    """Gated unit + residual / skip 1x1 convs: movenet/modules.py:49-93."""

    def __init__(self, residual_channels, skip_channels, dilation):
        super().__init__()
        self.dilation = dilation
        self.conv_filter = DilatedCausalConv1d(residual_channels, dilation=dilation)
        self.conv_gate = DilatedCausalConv1d(residual_channels, dilation=dilation)
        self.context_conv_filter = nn.Conv1d(residual_channels, residual_channels, 1)
        self.context_conv_gate = nn.Conv1d(residual_channels, residual_channels, 1)
        self.conv_residual = nn.Conv1d(residual_channels, residual_channels, 1)
        self.conv_skip = nn.Conv1d(residual_channels, skip_channels, 1)


class ResidualConvStack(_HotPathBlock):
    """``stack_size`` cycles of dilations 1, 2, ..., 2**(layer_size-1): movenet/modules.py:96-130."""

    def __init__(self, layer_size, stack_size, residual_channels, skip_channels):
        super().__init__()
        self.layer_size = layer_size
        self.stack_size = stack_size
        self.conv_layers = nn.ModuleList(
            GatedResidualConv1d(residual_channels, skip_channels, d) for d in self.dilations)

    @property
    def dilations(self) -> List[int]:
        return [1 << x for _ in range(self.stack_size) for x in range(self.layer_size)]


class DenseConv(_HotPathBlock):
    """Two 1x1 convs behind leaky-ReLUs: movenet/modules.py:133-142."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, 1)
        self.conv2 = nn.Conv1d(out_channels, out_channels, 1)
