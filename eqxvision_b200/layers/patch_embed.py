"""Image -> patch tokens (reference: layers/patch_embed.py:11-84).

conv PxP stride P (with bias) -> (D, gh, gw) -> ravel + moveaxis -> (gh*gw, D) -> norm.
On the device: a patch-row gather of the fp32 NCHW image followed by one tcgen05 GEMM
(`_engine.Plan._emit_patch_embed`).
"""
from typing import Optional, Tuple, Union

from .. import _trace as T
from .. import nn
from .. import random as jrandom


class PatchEmbed(nn.Module):
    img_size: Tuple[int]
    patch_size: Tuple[int]
    grid_size: Tuple[int]
    num_patches: int
    flatten: bool
    proj: nn.Conv2d
    norm: nn.Module

    def __init__(
        self,
        img_size: Union[int, Tuple[int]] = 224,
        patch_size: Union[int, Tuple[int]] = 16,
        in_chans: int = 3,
        embed_dim: int = 768,
        norm_layer=None,
        flatten: bool = True,
        *,
        key=None,
    ):
        self.img_size = img_size if isinstance(img_size, tuple) else (img_size, img_size)
        self.patch_size = patch_size if isinstance(patch_size, tuple) else (patch_size, patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        key = jrandom.PRNGKey(0) if key is None else key
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, key=key)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def __call__(self, x, *, key=None):
        _, h, w = x.shape
        if (h, w) != tuple(self.img_size):
            raise ValueError(f"Input image height ({h},{w}) doesn't match model ({self.img_size}).")
        x = self.proj(x)
        if self.flatten:
            x = T.to_tokens(x)  # (D, gh, gw) -> (gh*gw, D)
        return self.norm(x)
