"""Vision Transformer (reference: models/classification/vit.py, a DINO port).

Reference quirks that are kept on purpose: LayerNorm eps 1e-5, tanh-GELU,
`num_classes=0` by default (the output is the 768-d CLS feature), scale applied after q k^T, and
DINO/timm field order (cls_token, pos_embed, patch_embed, blocks, norm, fc) for positional loading.

Device lowering per block (A.2 of SURVEY.md): LayerNorm -> QKV GEMM(+bias) -> fused attention ->
proj GEMM(+bias+residual) -> LayerNorm -> fc1 GEMM(+bias+tanh-GELU) -> fc2 GEMM(+bias+residual).
The final LayerNorm is only evaluated on the CLS row that `x[0]` selects.
"""
from typing import Optional, Sequence, Tuple, Union

import torch

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import DropPath, MlpProjection, PatchEmbed
from ...utils import load_torch_weights


class _VitAttention(nn.Module):
    num_heads: int
    scale: float
    qkv: nn.Linear
    attn_drop: nn.Dropout
    proj: nn.Linear
    proj_drop: nn.Dropout

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, qk_scale=None,
                 attn_drop: float = 0.0, proj_drop: float = 0.0, *, key=None):
        k1, k2 = jrandom.split(key, 2)
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, use_bias=qkv_bias, key=k1)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, key=k2)
        self.proj_drop = nn.Dropout(proj_drop)

    def __call__(self, x, *, key=None):
        k1, k2 = jrandom.split(key, 2)
        # qkv columns are ordered (3, heads, head_dim): reshape(N,3,H,d) in vit.py:65
        out, attn = F.attention(self.qkv(x), self.num_heads, self.scale)
        attn = self.attn_drop(attn, key=k1)
        out = self.proj_drop(self.proj(out), key=k2)
        return out, attn


class _VitBlock(nn.Module):
    norm1: nn.Module
    attn: _VitAttention
    drop_path: DropPath
    norm2: nn.Module
    mlp: MlpProjection

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, act_layer=F.gelu, norm_layer=nn.LayerNorm, *, key):
        k1, k2 = jrandom.split(key, 2)
        self.norm1 = norm_layer(dim)
        self.attn = _VitAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                  attn_drop=attn_drop, proj_drop=drop, key=k1)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MlpProjection(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer,
                                 drop=drop, key=k2)

    def __call__(self, x, return_attention=False, *, key=None):
        ka, kd1, km, kd2 = jrandom.split(key, 4)
        y, attn = self.attn(self.norm1(x), key=ka)
        if return_attention:
            return attn
        x = x + self.drop_path(y, key=kd1)
        y = self.mlp(self.norm2(x), key=km)
        return x + self.drop_path(y, key=kd2)


class VisionTransformer(nn.Module):
    num_features: int
    cls_token: torch.Tensor
    pos_embed: torch.Tensor
    patch_embed: PatchEmbed
    pos_drop: nn.Dropout
    blocks: Sequence[_VitBlock]
    norm: nn.Module
    fc: nn.Linear
    inference: bool

    def __init__(
        self,
        img_size: Union[int, Tuple[int]] = 224,
        patch_size: Union[int, Tuple[int]] = 16,
        in_chans: int = 3,
        num_classes: int = 0,
        embed_dim: int = 768,
        depth: int = 12,
        num_heads: int = 12,
        mlp_ratio: float = 4.0,
        qkv_bias: bool = True,
        qk_scale=None,
        drop_rate=0.0,
        attn_drop_rate=0.0,
        drop_path_rate=0.0,
        norm_layer=nn.LayerNorm,
        *,
        key=None,
    ):
        key = jrandom.PRNGKey(0) if key is None else key
        keys = jrandom.split(key, depth + 3)
        self.inference = False
        self.num_features = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                      embed_dim=embed_dim)
        n_patches = self.patch_embed.num_patches
        self.cls_token = jrandom.truncated_normal(keys[0], -2, 2, (1, embed_dim))
        self.pos_embed = jrandom.truncated_normal(keys[1], -2, 2, (n_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        rates = [drop_path_rate * i / max(depth - 1, 1) for i in range(depth)]  # linspace(0, rate, depth)
        self.blocks = [
            _VitBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                      qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=rates[i],
                      norm_layer=norm_layer, key=keys[i + 1])
            for i in range(depth)
        ]
        self.norm = norm_layer(embed_dim)
        self.fc = nn.Identity() if num_classes == 0 else nn.Linear(embed_dim, num_classes, key=keys[-1])

    def _tokens(self, x):
        return F.prepend_cls_add_pos(self.patch_embed(x), self.cls_token, self.pos_embed)

    def __call__(self, x, *, key=None):
        keys = jrandom.split(key, len(self.blocks))
        x = self._tokens(x)
        for k, blk in zip(keys, self.blocks):
            x = blk(x, key=k)
        x = self.norm(x)          # row-wise; only row 0 survives the next line
        return self.fc(x[0])

    @nn.entrypoint
    def get_last_self_attention(self, x, *, key=None):
        if not self.inference:
            raise ValueError("Model being evaluated outside inference mode. Try in inference mode.")
        keys = jrandom.split(key, len(self.blocks))
        x = self._tokens(x)
        for k, blk in zip(keys[:-1], self.blocks[:-1]):
            x = blk(x, key=k)
        return self.blocks[-1](x, return_attention=True, key=key)


def _vit(patch_size, embed_dim, depth, num_heads, mlp_ratio, torch_weights, key, kwargs):
    model = VisionTransformer(patch_size=patch_size, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                              mlp_ratio=mlp_ratio, key=key, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def vit_tiny(patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, torch_weights: str = None, *,
             key=None, **kwargs):
    """ViT-Tiny/16 (vit.py:295-330)."""
    return _vit(patch_size, embed_dim, depth, num_heads, mlp_ratio, torch_weights, key, kwargs)


def vit_small(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, torch_weights: str = None, *,
              key=None, **kwargs):
    """ViT-Small/16 (vit.py:333-368)."""
    return _vit(patch_size, embed_dim, depth, num_heads, mlp_ratio, torch_weights, key, kwargs)


def vit_base(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, torch_weights: str = None, *,
             key=None, **kwargs):
    """ViT-Base/16 (vit.py:370-404)."""
    return _vit(patch_size, embed_dim, depth, num_heads, mlp_ratio, torch_weights, key, kwargs)
