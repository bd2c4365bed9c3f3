"""Swin Transformer v1 and v2 (reference: models/classification/swin.py, a torchvision port).

Field order follows the reference so that torchvision checkpoints load positionally
(`relative_position_bias_table`, `relative_position_index`, qkv, proj — the integer index buffer is
part of torchvision's state_dict and overwrites the leaf).

Reference quirks kept on purpose:
* `define_relative_position_index` (swin.py:314-335) discards its `jnp.stack(...)` and returns
  `ravel(relative_coords.sum(-1))` — partly negative indices (numpy wrap-around) until a checkpoint
  replaces them; `define_relative_position_bias_table` draws `truncated_normal(lower=2, upper=2)`.
* the feature map must be a multiple of the window (padding is commented out, swin.py:107-112).
* Swin-V2 (swin.py:369-522, 583-636): cosine attention whose q / k are normalised along AXIS 0 of the
  (num_windows, heads, tokens, d) arrays - over the windows of the image, not over d as in torchvision (swin.py:161-163);
  the continuous position bias MLP's channel-first output is reshaped to (-1, heads) without a transpose (swin.py:496-498),
  and its relative-position index has the same discarded-stack quirk as v1. All three are reproduced; the reference
  itself lists v2 pretrained weights as unsupported (docs/comparison.md:15).

Device lowering per block (activations stay channels-last, so every CHW<->HWC transpose of the
reference is free): LayerNorm -> qkv GEMM(+bias) -> eqxv_window_attention_bf16 (roll, window
partition, relative-position bias, shift mask, softmax, PV, reverse — all index arithmetic) ->
proj GEMM(+bias+residual) -> LayerNorm -> fc1 GEMM(+bias+tanh-GELU) -> fc2 GEMM(+bias+residual).
Patch merging = eqxv_patch_merge_bf16 gather -> LayerNorm(4C) -> GEMM(4C->2C)  (v2: GEMM, then LayerNorm(2C)).
V2 block: qkv GEMM (k bias zeroed) -> eqxv_swin_v2_qk_normalize_bf16 (in place) -> eqxv_window_attention_bf16 with
scale 1 and bias 16*sigmoid(cpb_mlp(coords)[index]) evaluated on the host (weights only) -> proj GEMM -> LayerNorm
-> + x (post-norm, swin.py:630-636).
"""
import math
import warnings
from functools import partial
from typing import Any, Callable, List, Optional

import torch

from ... import _trace as T
from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import DropPath, LayerNorm2d, Linear2d, MlpProjection
from ...utils import load_torch_weights


def _patch_merging_pad(x):
    """swin.py:23-33: x[:,0::2,0::2], x[:,1::2,0::2], x[:,0::2,1::2], x[:,1::2,1::2] stacked on channels"""
    return F.patch_merge(x)


def _get_relative_position_bias(relative_position_bias_table: torch.Tensor, relative_position_index: torch.Tensor,
                                window_size: List[int]) -> torch.Tensor:
    """swin.py:36-46 -> (heads, N, N); host-side, once per plan (parameters are constants in inference)"""
    n = window_size[0] * window_size[1]
    bias = relative_position_bias_table[relative_position_index.long()]  # negative indices wrap, as in numpy/jax
    return bias.reshape(n, n, -1).permute(2, 0, 1).contiguous()


class _PatchMerging(nn.Module):
    reduction: Linear2d
    norm: Callable

    def __init__(self, dim: int, norm_layer: Callable = LayerNorm2d, *, key=None):
        self.norm = norm_layer(4 * dim)
        self.reduction = Linear2d(4 * dim, 2 * dim, use_bias=False, key=key)

    def __call__(self, x, *, key=None):
        x = _patch_merging_pad(x)
        x = self.norm(x)
        return self.reduction(x)


class _PatchMergingV2(nn.Module):
    reduction: Linear2d
    norm: Callable

    def __init__(self, dim: int, norm_layer: Callable = LayerNorm2d, *, key=None):
        self.norm = norm_layer(2 * dim)
        self.reduction = Linear2d(4 * dim, 2 * dim, use_bias=False, key=key)

    def __call__(self, x, *, key=None):
        x = _patch_merging_pad(x)
        x = self.reduction(x)
        return self.norm(x)


def _shifted_window_attention(x, qkv: Linear2d, proj: Linear2d, relative_position_bias, window_size: List[int],
                              num_heads: int, shift_size: List[int], attention_dropout: float = 0.0,
                              dropout: float = 0.0, logit_scale=None, key=None):
    """swin.py:90-255 on a (C,H,W) map. Dropout with p == 0 is the identity (swin.py:17-20 divides by 1)."""
    if attention_dropout != 0.0 or dropout != 0.0:
        raise NotImplementedError("Swin dropout ignores inference mode in the reference (swin.py:227,233); "
                                  "only p == 0 is supported")
    c, h, w = x.shape
    shift = list(shift_size)
    if window_size[0] >= h:  # swin.py:115-119
        shift[0] = 0
    if window_size[1] >= w:
        shift[1] = 0
    tokens = F.to_tokens(x)                                   # transpose (1,2,0): free in NHWC
    head_dim = c // num_heads
    if logit_scale is not None:
        # Swin-V2 (swin.py:146-166): the k third of the qkv bias is zeroed, q and k are normalised (axis 0 = the windows
        # of the image) and q scaled by exp(min(logit_scale, log 100)) per head; no d^-1/2
        bias = qkv.bias
        if bias is not None:
            bias = bias.clone()
            length = bias.numel() // 3
            bias[length:2 * length] = 0
        qkv_t = T.linear(tokens, qkv.weight, bias)
        scale_q = torch.exp(torch.clamp(logit_scale.detach().float().reshape(-1), max=math.log(100.0)))
        out = F.window_attention(qkv_t, h, w, num_heads, window_size, shift, relative_position_bias, 1.0,
                                 cosine_scale=scale_q)
    else:
        qkv_t = nn.Linear.__call__(qkv, tokens)               # swin.py:155-157
        out = F.window_attention(qkv_t, h, w, num_heads, window_size, shift, relative_position_bias,
                                 head_dim ** -0.5)            # swin.py:168-231
    out = nn.Linear.__call__(proj, out)                       # swin.py:232
    return F.to_map(out, h, w)                                # reverse windows / roll / transpose: free


class _ShiftedWindowAttention(nn.Module):
    window_size: List[int]
    shift_size: List[int]
    num_heads: int
    attention_dropout: float
    dropout: float
    relative_position_bias_table: torch.Tensor
    relative_position_index: torch.Tensor
    qkv: nn.Linear
    proj: nn.Linear

    def __init__(self, dim: int, window_size: List[int], shift_size: List[int], num_heads: int,
                 qkv_bias: bool = True, proj_bias: bool = True, attention_dropout: float = 0.0,
                 dropout: float = 0.0, *, key=None):
        if len(window_size) != 2 or len(shift_size) != 2:
            raise ValueError("window_size and shift_size must be of length 2")
        keys = jrandom.split(key, 3)
        self.window_size = window_size
        self.shift_size = shift_size
        self.num_heads = num_heads
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.qkv = Linear2d(dim, dim * 3, use_bias=qkv_bias, key=keys[0])
        self.proj = Linear2d(dim, dim, use_bias=proj_bias, key=keys[1])
        self.relative_position_bias_table = self.define_relative_position_bias_table(key=keys[2])
        self.relative_position_index = self.define_relative_position_index()

    def define_relative_position_bias_table(self, key):
        shape = ((2 * self.window_size[0] - 1) * (2 * self.window_size[1] - 1), self.num_heads)
        return jrandom.truncated_normal(key, 2, 2, shape)     # degenerate bounds, as in swin.py:303-312

    def define_relative_position_index(self):
        ch = torch.arange(self.window_size[0])
        cw = torch.arange(self.window_size[1])
        coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij")).flatten(1)   # 2, Wh*Ww
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0)         # Wh*Ww, Wh*Ww, 2
        return rel.sum(-1).reshape(-1)                        # swin.py:334 (the stacked/offset version is discarded)

    def get_relative_position_bias(self) -> torch.Tensor:
        return _get_relative_position_bias(self.relative_position_bias_table, self.relative_position_index,
                                           self.window_size)

    def __call__(self, x, *, key=None):
        return _shifted_window_attention(
            x, self.qkv, self.proj, self.get_relative_position_bias(), self.window_size, self.num_heads,
            shift_size=self.shift_size, attention_dropout=self.attention_dropout, dropout=self.dropout, key=key)


class _ShiftedWindowAttentionV2(nn.Module):
    """swin.py:369-522"""
    window_size: List[int]
    shift_size: List[int]
    num_heads: int
    attention_dropout: float
    dropout: float
    logit_scale: torch.Tensor
    relative_position_bias_table: torch.Tensor
    relative_position_index: torch.Tensor
    qkv: nn.Linear
    proj: nn.Linear
    cpb_mlp: nn.Sequential

    def __init__(self, dim: int, window_size: List[int], shift_size: List[int], num_heads: int,
                 qkv_bias: bool = True, proj_bias: bool = True, attention_dropout: float = 0.0,
                 dropout: float = 0.0, *, key=None):
        if len(window_size) != 2 or len(shift_size) != 2:
            raise ValueError("window_size and shift_size must be of length 2")
        keys = jrandom.split(key, 3)
        self.window_size = window_size
        self.shift_size = shift_size
        self.num_heads = num_heads
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.qkv = Linear2d(dim, dim * 3, use_bias=qkv_bias, key=keys[0])
        self.proj = Linear2d(dim, dim, use_bias=proj_bias, key=keys[1])
        self.relative_position_bias_table = self.define_relative_position_bias_table(key=keys[2])
        self.relative_position_index = self.define_relative_position_index()
        self.logit_scale = torch.log(10 * torch.ones((num_heads, 1, 1)))
        # mlp to generate continuous relative position bias (swin.py:413-421); evaluated on the host, weights only
        self.cpb_mlp = nn.Sequential([
            nn.Lambda(_chw_of_hwc),
            Linear2d(2, 512, use_bias=True, key=keys[1]),
            nn.Lambda(F.relu),
            Linear2d(512, num_heads, use_bias=False, key=keys[2]),
            nn.Lambda(F.identity),
        ])
        if qkv_bias:                                          # swin.py:422-428
            length = self.qkv.bias.numel() // 3
            self.qkv.bias[length:2 * length] = 0

    define_relative_position_index = _ShiftedWindowAttention.define_relative_position_index   # swin.py:430-452

    def define_relative_position_bias_table(self, key):
        """swin.py:454-481: the log-spaced relative COORDINATE table (2Wh-1, 2Ww-1, 2), input of cpb_mlp"""
        wh, ww = self.window_size
        ch = torch.arange(-(wh - 1), wh, dtype=torch.float32)
        cw = torch.arange(-(ww - 1), ww, dtype=torch.float32)
        t = torch.stack(torch.meshgrid(ch, cw, indexing="ij")).permute(1, 2, 0)
        t = torch.stack([t[:, :, 0] / wh - 1, t[:, :, 1] / ww - 1], -1)   # (sic) "/ window - 1", swin.py:468-469
        t = 8 * t
        return torch.sign(t) * torch.log2(torch.abs(t) + 1.0) / 3.0

    def get_relative_position_bias(self) -> torch.Tensor:
        """swin.py:483-504 on the host (parameters only): 16 * sigmoid(cpb_mlp(table).reshape(-1, heads)[index])"""
        table = self.relative_position_bias_table.detach().float()
        a, b, _ = table.shape
        fc1, fc2 = self.cpb_mlp.layers[1], self.cpb_mlp.layers[3]
        tok = table.reshape(a * b, 2)                          # Linear2d sees the (2, A, B) map as (A*B, 2) rows
        hid = torch.relu(tok @ fc1.weight.detach().float().t() + fc1.bias.detach().float())
        out_chw = (hid @ fc2.weight.detach().float().t()).t().contiguous()   # Linear2d returns (heads, A, B)
        flat = out_chw.reshape(-1, self.num_heads)             # (sic) channel-first data read as (-1, heads)
        bias = _get_relative_position_bias(flat, self.relative_position_index, self.window_size)
        return 16 * torch.sigmoid(bias)

    def __call__(self, x, *, key=None):
        return _shifted_window_attention(
            x, self.qkv, self.proj, self.get_relative_position_bias(), self.window_size, self.num_heads,
            shift_size=self.shift_size, attention_dropout=self.attention_dropout, dropout=self.dropout,
            logit_scale=self.logit_scale, key=key)


def _chw_of_hwc(x, *, key=None):   # Lambda(partial(jnp.transpose, axes=(2, 0, 1))), swin.py:415 (host-side only)
    return x.permute(2, 0, 1)


class _SwinTransformerBlock(nn.Module):
    norm1: Callable
    attn: nn.Module
    stochastic_depth: DropPath
    norm2: Callable
    mlp: MlpProjection

    def __init__(self, dim: int, num_heads: int, window_size: List[int], shift_size: List[int],
                 mlp_ratio: float = 4.0, dropout: float = 0.0, attention_dropout: float = 0.0,
                 stochastic_depth_prob: float = 0.0, norm_layer: Callable = LayerNorm2d,
                 attn_layer: Callable = _ShiftedWindowAttention, *, key=None):
        keys = jrandom.split(key, 2)
        self.norm1 = norm_layer(dim)
        self.attn = attn_layer(dim, window_size, shift_size, num_heads, attention_dropout=attention_dropout,
                               dropout=dropout, key=keys[0])
        self.stochastic_depth = DropPath(stochastic_depth_prob, mode="local")
        self.norm2 = norm_layer(dim)
        self.mlp = MlpProjection(dim, int(dim * mlp_ratio), dim, lin_layer=Linear2d, act_layer=F.gelu,
                                 drop=dropout, key=keys[1])

    def __call__(self, x, *, key=None):
        keys = jrandom.split(key, 4)
        x = x + self.stochastic_depth(self.attn(self.norm1(x), key=keys[0]), key=keys[1])
        x = x + self.stochastic_depth(self.mlp(self.norm2(x), key=keys[2]), key=keys[3])
        return x


class _SwinTransformerBlockV2(_SwinTransformerBlock):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0, dropout=0.0,
                 attention_dropout=0.0, stochastic_depth_prob=0.0, norm_layer=LayerNorm2d,
                 attn_layer=_ShiftedWindowAttentionV2, *, key=None):
        super().__init__(dim, num_heads, window_size, shift_size, mlp_ratio, dropout, attention_dropout,
                         stochastic_depth_prob, norm_layer, attn_layer, key=key)

    def __call__(self, x, *, key=None):  # post-norm, swin.py:630-636
        keys = jrandom.split(key, 4)
        x = x + self.stochastic_depth(self.norm1(self.attn(x, key=keys[0])), key=keys[1])
        x = x + self.stochastic_depth(self.norm2(self.mlp(x, key=keys[2])), key=keys[3])
        return x


class SwinTransformer(nn.Module):
    """swin.py:639-772"""
    features: nn.Sequential
    norm: Callable
    avgpool: nn.AdaptiveAvgPool2d
    head: nn.Linear

    def __init__(self, patch_size: List[int], embed_dim: int, depths: List[int], num_heads: List[int],
                 window_size: List[int], mlp_ratio: float = 4.0, dropout: float = 0.0,
                 attention_dropout: float = 0.0, stochastic_depth_prob: float = 0.1, num_classes: int = 1000,
                 norm_layer: Callable = None, block: Optional[Callable] = None,
                 downsample_layer: Optional[Callable] = None, *, key=None):
        if key is None:
            key = jrandom.PRNGKey(0)
        keys = jrandom.split(key, 2)
        if block is None:
            block = _SwinTransformerBlock
        if norm_layer is None:
            norm_layer = partial(LayerNorm2d, eps=1e-5)
        if downsample_layer is None:
            downsample_layer = _PatchMerging

        layers: List[nn.Module] = [nn.Sequential([
            nn.Conv2d(3, embed_dim, kernel_size=(patch_size[0], patch_size[1]),
                      stride=(patch_size[0], patch_size[1]), key=keys[0]),
            norm_layer(embed_dim),
        ])]
        total_stage_blocks = sum(depths)
        stage_block_id = 0
        for i_stage in range(len(depths)):
            stage: List[nn.Module] = []
            dim = embed_dim * 2 ** i_stage
            for i_layer in range(depths[i_stage]):
                keys = jrandom.split(keys[1], 2)
                sd_prob = stochastic_depth_prob * float(stage_block_id) / (total_stage_blocks - 1)
                stage.append(block(dim, num_heads[i_stage], window_size=window_size,
                                   shift_size=[0 if i_layer % 2 == 0 else w // 2 for w in window_size],
                                   mlp_ratio=mlp_ratio, dropout=dropout, attention_dropout=attention_dropout,
                                   stochastic_depth_prob=sd_prob, norm_layer=norm_layer, key=keys[0]))
                stage_block_id += 1
            layers.append(nn.Sequential(stage))
            if i_stage < len(depths) - 1:
                keys = jrandom.split(keys[1], 2)
                layers.append(downsample_layer(dim, norm_layer, key=keys[0]))
        self.features = nn.Sequential(layers)
        num_features = embed_dim * 2 ** (len(depths) - 1)
        self.norm = norm_layer(num_features)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.head = nn.Linear(num_features, num_classes, key=keys[1])

    def __call__(self, x, *, key=None):
        keys = jrandom.split(key, 2)
        x = self.features(x, key=keys[0])
        x = self.norm(x)
        x = self.avgpool(x)
        x = F.ravel(x)
        return self.head(x, key=keys[1])


def _swin_transformer(arch: str, patch_size, embed_dim, depths, num_heads, window_size, stochastic_depth_prob,
                      torch_weights, **kwargs: Any) -> SwinTransformer:
    warnings.warn("Currently, dynamic padding of the input is not supported! "
                  "Please make sure that the input is a multiple of window_size.")
    model = SwinTransformer(patch_size=patch_size, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
                            window_size=window_size, stochastic_depth_prob=stochastic_depth_prob, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def swin_t(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:806-826"""
    return _swin_transformer("swin_t", [4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [7, 7], 0.2, torch_weights,
                             **kwargs)


def swin_s(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:829-848"""
    return _swin_transformer("swin_s", [4, 4], 96, [2, 2, 18, 2], [3, 6, 12, 24], [7, 7], 0.3, torch_weights,
                             **kwargs)


def swin_b(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:851-871"""
    return _swin_transformer("swin_b", [4, 4], 128, [2, 2, 18, 2], [4, 8, 16, 32], [7, 7], 0.5, torch_weights,
                             **kwargs)


def swin_v2_t(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:874-896"""
    return _swin_transformer("swin_v2_t", [4, 4], 96, [2, 2, 6, 2], [3, 6, 12, 24], [8, 8], 0.2, torch_weights,
                             block=_SwinTransformerBlockV2, downsample_layer=_PatchMergingV2, **kwargs)


def swin_v2_s(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:899-921"""
    return _swin_transformer("swin_v2_s", [4, 4], 96, [2, 2, 18, 2], [3, 6, 12, 24], [8, 8], 0.3, torch_weights,
                             block=_SwinTransformerBlockV2, downsample_layer=_PatchMergingV2, **kwargs)


def swin_v2_b(torch_weights: str = None, **kwargs: Any) -> SwinTransformer:
    """swin.py:924-946"""
    return _swin_transformer("swin_v2_b", [4, 4], 128, [2, 2, 18, 2], [4, 8, 16, 32], [8, 8], 0.5, torch_weights,
                             block=_SwinTransformerBlockV2, downsample_layer=_PatchMergingV2, **kwargs)
