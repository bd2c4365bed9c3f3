"""Conv -> norm -> activation block (reference: layers/conv_norm_activation.py:10-86).

The three layers are kept as separate sequence entries (the positional weight loader walks
conv.weight, [conv.bias], bn.weight, bn.bias in that order), but on the device they are ONE
kernel: the tracer folds the BatchNorm affine into the filter/bias and the activation into the
implicit-GEMM epilogue (csrc/igemm.cu).
"""
from functools import partial
from typing import Callable, Optional

from .. import functional as F
from .. import nn
from .. import random as jrandom


class ConvNormActivation(nn.Sequential):
    out_channels: int

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        kernel_size: int = 3,
        stride: int = 1,
        padding: Optional[int] = None,
        groups: int = 1,
        norm_layer: Optional[Callable] = nn.BatchNorm,
        activation_layer: Optional[Callable] = F.relu,
        dilation: int = 1,
        use_bias: Optional[bool] = None,
        *,
        key=None,
    ) -> None:
        key = jrandom.PRNGKey(0) if key is None else key
        pad = (kernel_size - 1) // 2 * dilation if padding is None else padding
        bias = (norm_layer is None) if use_bias is None else use_bias
        seq = [nn.Conv2d(in_channels, out_channels, kernel_size, stride, pad, dilation=dilation,
                         groups=groups, use_bias=bias, key=key)]
        if norm_layer is not None:
            base = norm_layer.func if isinstance(norm_layer, partial) else norm_layer
            if base is nn.BatchNorm:
                seq.append(norm_layer(out_channels, axis_name="batch"))
            else:
                seq.append(norm_layer(out_channels))
        if activation_layer is not None:
            seq.append(nn.Lambda(activation_layer))
        super().__init__(seq)
        self.out_channels = out_channels
