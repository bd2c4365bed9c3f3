"""Channel-wise LayerNorm / Linear on (C,H,W) maps (reference: layers/extensions_2d.py:9-50).

The reference transposes (C,HW)->(HW,C), vmaps the 1-D layer and transposes back. Activations
here are already channels-last, so both transposes are free re-interpretations of the same buffer.
"""
from .. import _trace as T
from .. import nn


class LayerNorm2d(nn.LayerNorm):
    def __call__(self, x, *, key=None):
        c, h, w = x.shape
        y = T.layer_norm(T.to_tokens(x), self.weight, self.bias, self.eps)
        return T.to_map(y, h, w)


class Linear2d(nn.Linear):
    def __call__(self, x, *, key=None):
        c, h, w = x.shape
        y = T.linear(T.to_tokens(x), self.weight, self.bias)
        return T.to_map(y, h, w)
