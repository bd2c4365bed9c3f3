"""Plan builder and executor: the replacement for `eqx.filter_jit(jax.vmap(net, axis_name="batch"))`.

A model's per-sample `__call__` is traced once on a `Sym` (see `_trace.py`); the resulting
expression DAG is lowered here to a flat list of C-ABI kernel launches over preallocated
channels-last bf16 buffers (batch dimension folded in), which is then captured into ONE CUDA graph
per (model, entry point, batch, input shape). Replaying the graph is the whole forward pass: no
Python, no allocation, no host synchronisation inside.

Data layout in HBM: every activation is a row-major matrix [rows, pitch] of bf16 with
rows = N*H*W (feature maps, i.e. NHWC) or N*T (token matrices) or N (vectors) and
pitch = channels rounded up to 8 (pad columns are zero and stay zero).
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, _pack, ops
from . import _trace as T
from . import random as jrandom
from ._lib import ACT_BY_NAME, EqxvError

BF16 = torch.bfloat16


def _round8(v: int) -> int:
    return (v + 7) // 8 * 8


# The tcgen05 GEMM stages the fp32 shift of ALL its output channels in shared memory (20 KB: csrc/igemm.cu), i.e.
# at most ~5000 output channels per launch. Wider layers (ConvNeXt-Large MLP 6144, RegNetY-128GF 7392) are lowered as
# several launches over column slices of the same output buffer; nothing else changes (same A operand, same epilogue).
MAX_COUT_PER_LAUNCH = 4096


def _n_chunks(cout: int, limit: Optional[int] = None):
    limit = MAX_COUT_PER_LAUNCH if limit is None else limit
    if cout <= limit:
        return [(0, cout)]
    k = -(-cout // limit)
    per = _round8(-(-cout // k))
    return [(n0, min(cout, n0 + per)) for n0 in range(0, cout, per)]


class Buf:
    """A device activation: 2-D view [rows, c] with row stride `pitch` (a torch tensor as holder)."""
    __slots__ = ("t", "c", "n", "geom")

    def __init__(self, t: torch.Tensor, c: int, n: int, geom: Tuple[int, ...]):
        self.t, self.c, self.n, self.geom = t, c, n, geom  # geom: (h, w) for maps, (tokens,) or ()

    @property
    def pitch(self) -> int:
        return self.t.stride(0)

    @property
    def cpad(self) -> int:
        return _round8(self.c)

    def rows(self, cols: Optional[int] = None) -> torch.Tensor:
        cols = self.cpad if cols is None else cols
        return torch.as_strided(self.t, (self.t.shape[0], cols), (self.pitch, 1), self.t.storage_offset())

    def map(self, h: int, w: int, cols: Optional[int] = None) -> torch.Tensor:
        cols = self.cpad if cols is None else cols
        p = self.pitch
        assert self.t.shape[0] == self.n * h * w
        return torch.as_strided(self.t, (self.n, h, w, cols), (h * w * p, w * p, p, 1), self.t.storage_offset())


class Plan:
    def __init__(self, device: torch.device, batch: int, in_shape: Tuple[int, ...], u8: Optional[dict] = None):
        """`u8`: None for the reference's fp32 NCHW input, or dict(mean, std, raw_hw) for the uint8 HWC input edge
        (transforms.ImagesU8): ToTensor + Normalize (+ Resize) then run on the device, fused into the first layer's
        layout kernel (csrc/input_edge.cu)."""
        self.device = device
        self.n = batch
        self.in_shape = tuple(in_shape)
        self.steps: List[Tuple[Callable, dict]] = []
        self.consts: List[torch.Tensor] = []  # keeps packed weights alive
        self.memo: Dict[int, Buf] = {}
        self.keep: List[Any] = []             # keeps traced exprs alive (ids are memo keys)
        self.u8 = u8
        self._arena_candidates: Dict[int, torch.Tensor] = {}
        self._ln_producers: Dict[int, int] = {}     # id of a Linear expr -> index of its plain residual-GEMM step
        self._ln_stats: Dict[int, torch.Tensor] = {}  # id of that expr -> row statistics its patched step writes
        self._pool_requests: Dict[int, Buf] = {}   # id of a depthwise Conv expr -> pooled Buf: squeeze fused into that conv
        self.arena = None
        self.stream: Optional[int] = None     # lane stream (created by build_plan on a CUDA device)
        self.done = None                      # event: the lane's last graph launch has finished
        self.out_ready = None                 # event: the lane's last outputs have been copied out
        if u8 is None:
            self.x_in = torch.empty((batch,) + self.in_shape, dtype=torch.float32, device=device)
            self.x_host_target = self.x_in
        else:
            from . import transforms

            if len(self.in_shape) != 3:
                raise EqxvError("uint8 image batches feed models with a (C, H, W) per-sample input")
            c, h, w = self.in_shape
            self.x_in = torch.empty((batch, h, w, c), dtype=torch.uint8, device=device)
            self.lut = self.const(transforms.normalize_lut(u8["mean"], u8["std"]))
            rh, rw = u8.get("raw_hw") or (h, w)
            if (rh, rw) != (h, w):
                # transforms.Resize of the reference fixture (tests/conftest.py:25), on the device
                self.x_host_target = torch.empty((batch, rh, rw, c), dtype=torch.uint8, device=device)
                self.step(ops.u8_resize_bilinear, x=self.x_host_target, oh=h, ow=w, out=self.x_in)
            else:
                self.x_host_target = self.x_in
        self._input_nhwc: Dict[int, Buf] = {}
        self._input_stem: Dict[int, torch.Tensor] = {}
        self._concat_groups: Dict[int, dict] = {}
        self.outputs: List[Tuple[torch.Tensor, Tuple[int, ...]]] = []
        self.out_struct = None
        self.graph = None
        self.act_bytes = 0

    # -------------------------------------------------------------- allocation / constants
    def alloc(self, rows: int, c: int, geom=(), dtype=BF16) -> Buf:
        pitch = _round8(c)
        t = torch.zeros((rows, pitch), dtype=dtype, device=self.device)
        self.act_bytes += t.numel() * t.element_size()
        if pitch == c:
            self.arena_ok(t)   # every column is rewritten by the producer on each replay: the memory can be shared
        return Buf(t, c, self.n, tuple(geom))

    def arena_ok(self, t: torch.Tensor) -> torch.Tensor:
        """mark a buffer as fully overwritten by its producer(s) on every replay (no zero-filled pad columns or concat
        slots to preserve): `plan_memory` may then place it in the shared arena"""
        self._arena_candidates[t.untyped_storage().data_ptr()] = t
        return t

    def arena_pin(self, t: torch.Tensor) -> None:
        """keep a buffer out of the arena after all (several producers fill disjoint slices over time)"""
        self._arena_candidates.pop(t.untyped_storage().data_ptr(), None)

    def plan_memory(self, align: int = 1024) -> None:
        """Liveness-based activation arena. Every shareable buffer lives from the first step that touches it to the last
        (outputs: to the end of the replay); buffers whose lifetimes do not overlap share addresses inside ONE
        allocation (first-fit over a free list, in order of first use). Steps, outputs and views are then re-pointed
        into the arena and the individual allocations are released. Ordering between consecutive launches is what makes
        this safe: every kernel of the library waits for its predecessor's completion (griddepcontrol.wait) before it
        touches activation memory. ResNet-50 at batch 256: 5.4 GiB -> ~1 GiB."""
        cands = self._arena_candidates
        if not cands or os.environ.get("EQXV_NO_ARENA") == "1":
            return
        first: Dict[int, int] = {}
        last: Dict[int, int] = {}

        def tensors_of(kw):
            for v in kw.values():
                if isinstance(v, torch.Tensor):
                    yield v

        for i, (_, kw) in enumerate(self.steps):
            for t in tensors_of(kw):
                p_ = t.untyped_storage().data_ptr()
                if p_ in cands:
                    first.setdefault(p_, i)
                    last[p_] = i
        end = len(self.steps)
        for o, _ in self.outputs:
            p_ = o.untyped_storage().data_ptr()
            if p_ in cands and p_ in first:
                last[p_] = end
        order = sorted(first, key=lambda p_: (first[p_], -cands[p_].untyped_storage().nbytes()))
        live: List[Tuple[int, int, int]] = []      # (last step, offset, size) of placed buffers still alive
        free: List[Tuple[int, int]] = []           # (offset, size) holes below `top`
        top = peak = 0
        offset: Dict[int, int] = {}
        for p_ in order:
            now = first[p_]
            still = []
            for l_, off, size in live:
                if l_ < now:
                    free.append((off, size))
                else:
                    still.append((l_, off, size))
            live = still
            free.sort()
            merged: List[Tuple[int, int]] = []
            for off, size in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + size)
                else:
                    merged.append((off, size))
            if merged and merged[-1][0] + merged[-1][1] == top:   # a hole at the top shrinks the arena front
                top = merged.pop()[0]
            free = merged
            need = -(-cands[p_].untyped_storage().nbytes() // align) * align
            best = None
            for j, (off, size) in enumerate(free):
                if size >= need and (best is None or size < free[best][1]):
                    best = j
            if best is not None:
                off, size = free.pop(best)
                if size > need:
                    free.append((off + need, size - need))
            else:
                off = top
                top += need
            offset[p_] = off
            peak = max(peak, off + need)
            live.append((last[p_], off, need))
        total = peak
        arena = torch.empty(max(total, align), dtype=torch.uint8, device=self.device)
        ast = arena.untyped_storage()

        def remap(t):
            if not isinstance(t, torch.Tensor):
                return t
            off = offset.get(t.untyped_storage().data_ptr())
            if off is None:
                return t
            es = t.element_size()
            return torch.empty(0, dtype=t.dtype, device=t.device).set_(ast, off // es + t.storage_offset(), t.shape,
                                                                       t.stride())

        self.steps = [(fn, {k: remap(v) for k, v in kw.items()}) for fn, kw in self.steps]
        self.outputs = [(remap(o), shp) for o, shp in self.outputs]
        shared = sum(cands[p_].untyped_storage().nbytes() for p_ in offset)
        self.act_bytes_unshared = self.act_bytes
        self.act_bytes = self.act_bytes - shared + total
        self.arena = arena
        self._arena_candidates = {}
        self.memo.clear()
        self._input_nhwc.clear()
        self._input_stem.clear()
        self._concat_groups.clear()

    def const(self, t: torch.Tensor) -> torch.Tensor:
        d = t.to(self.device).contiguous()
        self.consts.append(d)
        return d

    def step(self, fn: Callable, **kw):
        self.steps.append((fn, kw))

    # -------------------------------------------------------------- emission
    _DIRECT_DST = (T.Conv, T.Pool, T.Resize, T.ChannelView)  # nodes that can write straight into a caller-provided slice

    def emit(self, sym: T.Sym, out_f32: bool = False, dst: Optional[Buf] = None) -> Buf:
        """Lower `sym` (memoised). With `dst` the result must end up in that buffer slice: producers
        that support it store there directly, anything else is computed and copied."""
        key = id(sym.expr)
        if key in self.memo and not out_f32:
            buf = self.memo[key]
            if dst is not None and buf.t.data_ptr() != dst.t.data_ptr():
                self.step(ops.copy2d, dst=dst.rows(), src=buf.rows(dst.cpad))
            return buf
        e = sym.expr
        fn = getattr(self, "_emit_" + type(e).__name__, None)
        if fn is None:
            raise NotImplementedError(f"no lowering for {type(e).__name__}")
        self.keep.append(e)
        if isinstance(e, T.Conv):
            buf = fn(sym, e, out_f32, dst)
        elif isinstance(e, T.Linear):
            buf = fn(sym, e, out_f32)
        elif dst is not None and isinstance(e, self._DIRECT_DST):
            buf = fn(sym, e, dst)
        else:
            buf = fn(sym, e)
            if dst is not None:
                self.step(ops.copy2d, dst=dst.rows(), src=buf.rows(dst.cpad))
                buf = dst
        if not out_f32:
            self.memo[key] = buf
        return buf

    # ---- input -------------------------------------------------------------------------------
    def _emit_Input(self, sym, e):
        raise EqxvError("the raw fp32 input can only feed a convolution / patch embedding")

    def x_f32(self) -> torch.Tensor:
        """the fp32 NCHW model input: the caller's batch, or Normalize(ToTensor(pixels)) for uint8 plans (only the
        layouts without a fused uint8 kernel come here)"""
        if self.u8 is None:
            return self.x_in
        if not hasattr(self, "_x_f32"):
            self._x_f32 = torch.empty((self.n,) + self.in_shape, dtype=torch.float32, device=self.device)
            self.step(ops.u8_to_nchw_f32, x=self.x_in, lut=self.lut, out=self._x_f32)
        return self._x_f32

    def input_nhwc(self, c_pad: int) -> Buf:
        if c_pad not in self._input_nhwc:
            c, h, w = self.in_shape
            buf = self.alloc(self.n * h * w, c_pad, (h, w))
            if self.u8 is not None and c_pad == 8:
                self.step(ops.u8_to_nhwc, x=self.x_in, lut=self.lut, out=buf.map(h, w, c_pad))
            else:
                self.step(ops.nchw_to_nhwc, x=self.x_f32(), c_pad=c_pad, out=buf.map(h, w, c_pad))
            buf.c = c
            self._input_nhwc[c_pad] = buf
        return self._input_nhwc[c_pad]

    STEM_C4 = os.environ.get("EQXV_NO_STEM_C4") != "1"   # pixel-pair layout for stride-2 stems (A/B switch)

    def _stem_input(self, ph: int, h: int, wd: int, c4: bool = False) -> torch.Tensor:
        """the padded image the first-layer kernels read (eqxv_pack_stem_input[_c4] / their uint8 twins), once per padding
        and layout: 8 channels per pixel, or - c4 - pairs of 4-channel pixels"""
        key = (ph, c4)
        if key not in self._input_stem:
            # the pack kernels write the zero border too: the whole buffer is rewritten every replay
            shape = (self.n, h + 2 * ph, (wd + 8) // 2, 8) if c4 else (self.n, h + 2 * ph, wd + 8, 8)
            xp = self.arena_ok(torch.zeros(shape, dtype=BF16, device=self.device))
            self.act_bytes += xp.numel() * 2
            if self.u8 is not None:
                self.step(ops.u8_pack_stem_input_c4 if c4 else ops.u8_pack_stem_input, x=self.x_in, lut=self.lut, pad=ph, out=xp)
            elif c4:
                self.step(ops.pack_stem_input_c4, x_nchw=self.x_in, pad=ph, out=xp)
            else:
                self.step(ops.pack_stem_input, x_nchw=self.x_in, pad=ph, out=xp)
            self._input_stem[key] = xp
        return self._input_stem[key]

    # ---- conv --------------------------------------------------------------------------------
    @staticmethod
    def _epilogue(e) -> Tuple[int, bool]:
        if e.act1 is not None and e.res is not None and e.act2 is not None:
            raise EqxvError("internal: unfusable epilogue")
        if e.res is not None and e.act1 is not None:
            return ACT_BY_NAME[e.act1], True
        return ACT_BY_NAME[e.act2 if e.res is not None else e.act1], False

    def _emit_Conv(self, sym, e: T.Conv, out_f32=False, dst: Optional[Buf] = None):
        cout, cin_g, kh, kw = e.weight.shape
        (sh, sw), (ph, pw), (dh, dw) = e.stride, e.padding, e.dilation
        ho, wo = sym.shape[1:]
        if dst is None and not out_f32 and self.BNECK_FUSE:
            # before the shortcut operand is lowered: the fused kernel wants to lower the trunk first (see there)
            fused = self._emit_bottleneck64(sym, e)
            if fused is not None:
                return fused
        w, b = _pack.fold_bn(e.weight, e.bias, e.bn)
        act, res_after = self._epilogue(e)
        if e.res is not None and self.BNECK_FUSE and isinstance(e.x.expr, T.Conv) and e.groups == 1:
            # trunk before shortcut: the trunk's first convolution may come out of the launch that produces the shortcut
            # tensor (fused bottleneck + next 1x1); lowering the shortcut first would emit that block without it
            self.emit(e.x)
        res = self.emit(e.res) if e.res is not None else None
        xin = e.x
        c_in, h, wd = xin.shape

        if e.groups != 1:
            if e.groups == c_in and cin_g == 1 and cout == c_in:
                return self._emit_depthwise(sym, e, w, b, act, res, dst)
            return self._emit_grouped(sym, e, w, b, act, res_after, res, dst, out_f32)
        if sh != sw or ph != pw or dh != dw:
            raise NotImplementedError("anisotropic stride/padding/dilation is not supported")

        bias_d = self.const(b) if b is not None else None
        if dst is not None:
            if dst.c != cout or out_f32:
                raise EqxvError("internal: destination slice does not match the convolution output")
            out = dst
        else:
            out = self.alloc(self.n * ho * wo, cout, (ho, wo), dtype=torch.float32 if out_f32 else BF16)

        if (self.GATE_FUSE and isinstance(xin.expr, T.ChannelScale) and id(xin.expr) not in self.memo and kh == kw == 1
                and sh == 1 and ph == 0 and act == _lib.ACT_NONE and not res_after and not out_f32 and bias_d is not None
                and cout <= MAX_COUT_PER_LAUNCH):
            # SqueezeExcitation gate feeding the projection (efficientnet.py:161-170): applied to the GEMM's A operand
            # in shared memory (eqxv_gemm_gated_bf16) instead of a pass that reads and rewrites the expanded tensor
            ce = xin.expr
            sb = self.emit(ce.s)          # gate first: its squeeze may ride in the depthwise kernel
            xb = self.emit(ce.x)
            if xb.pitch == c_in and c_in % 8 == 0 and sb.pitch >= c_in:
                tail = self.TAIL_SHIFT and _pack.tail_shift_applies(c_in)
                wp = self.const(_pack.pack_conv_weight(w, c_in, tail_shift=tail))
                self.step(ops.gemm_gated, a=xb.rows(c_in), gate=sb.rows(sb.pitch), wgt=wp, bias=bias_d,
                          rows_per_image=h * wd, residual=None if res is None else res.rows(), out=out.rows(),
                          k_tail_shift=tail)
                return out
        if isinstance(xin.expr, T.Input):
            if c_in <= 8 and kh <= 8 and kw <= 8 and sh in (1, 2, 4) and dh == 1 and 2 * ph <= kw \
                    and res is None and not out_f32 and dst is None:
                # first-layer conv on the raw image (resnet.py:243-251, vgg.py:137, efficientnet.py:327):
                # padded NHWC8 image + one GEMM K-block per filter row (eqxv_conv_stem_bf16)
                # stride-2 stems with <= 4 channels on an even width: pixel-pair layout (half the MMAs / staged bytes)
                c4 = self.STEM_C4 and sh == 2 and c_in <= 4 and wd % 2 == 0 and cout <= 256
                wp = self.const(_pack.pack_stem_weight_c4(w) if c4 else _pack.pack_stem_weight(w))
                self.step(ops.conv_stem, xpad=self._stem_input(ph, h, wd, c4), wgt=wp, bias=bias_d, n=self.n, h=h, w=wd, c4=c4,
                          cout=cout, kh=kh, kw=kw, stride=sh, pad=ph, act=act, out=out.map(ho, wo))
                return out
            xb = self.input_nhwc(_round8(c_in))
        else:
            xb = self.emit(xin)
        cin_eff = xb.cpad
        # every K chunk inside the tensor (EQXV_FLAG_K_TAIL_SHIFT): partially out-of-bounds TMA boxes are slow
        tail = self.TAIL_SHIFT and _pack.tail_shift_applies(cin_eff)
        wp = self.const(_pack.pack_conv_weight(w, cin_eff, tail_shift=tail))
        omap, rmap = out.map(ho, wo, cout if out_f32 else None), None if res is None else res.map(ho, wo)
        for n0, n1 in _n_chunks(cout):
            whole = (n0, n1) == (0, cout)
            self.step(ops.conv2d, x=xb.map(h, wd, cin_eff), wgt=wp[n0:n1], bias=None if bias_d is None else bias_d[n0:n1],
                      cin=cin_eff, cout=n1 - n0, kh=kh, kw=kw, stride=sh, pad=ph, dil=dh, act=act,
                      residual=None if rmap is None else rmap[..., n0:n1], res_after_act=res_after,
                      out=omap if whole else omap[..., n0:n1], out_f32=out_f32, k_tail_shift=tail)
        return out

    # ---- fused ResNet bottleneck (64-channel trunk) ---------------------------------------------
    BNECK_FUSE = os.environ.get("EQXV_NO_BNECK") != "1"
    TAIL_SHIFT = os.environ.get("EQXV_NO_TAIL_SHIFT") != "1"   # in-bounds last K chunk (A/B switch)

    def _plain_conv(self, e, cin, cout, k, pad, act_name, res: bool) -> bool:
        """e is a dense stride-1 undilated k x k convolution cin -> cout with exactly this epilogue and a bias"""
        if not isinstance(e, T.Conv) or e.groups != 1 or tuple(e.weight.shape) != (cout, cin, k, k):
            return False
        if e.stride != (1, 1) or e.padding != (pad, pad) or e.dilation != (1, 1) or (e.bn is None and e.bias is None):
            return False
        if res:
            return e.res is not None and e.act1 is None and e.act2 == act_name
        return e.res is None and e.act2 is None and e.act1 == act_name

    def _bottleneck64_match(self, e):
        """e = the closing 1x1 convolution (64 -> 256, + shortcut, ReLU) of a bottleneck whose 3x3 is 64 -> 64, stride 1
        (resnet.py:144-162 in layer1 of ResNet-50/101/152) and has not been lowered yet"""
        if id(e) in self.memo or not self._plain_conv(e, 64, 256, 1, 0, "relu", True):
            return None
        c2 = e.x.expr
        if id(c2) in self.memo or not self._plain_conv(c2, 64, 64, 3, 1, "relu", False):
            return None
        if e.res.shape != (256,) + tuple(e.x.shape[1:]) or isinstance(c2.x.expr, T.Input):
            return None
        down = e.res.expr
        if id(down) in self.memo or not self._plain_conv(down, 64, 256, 1, 0, None, False) or \
                isinstance(down.x.expr, T.Input):
            down = None
        return c2, down

    def _emit_bottleneck64(self, sym, e: T.Conv) -> Optional[Buf]:
        """Lowers a matched bottleneck onto eqxv_bottleneck64_fused_bf16. Called for the block's closing convolution
        (returns y) or for the NEXT block's opening 1x1 (256 -> 64 | 128, ReLU) whose input is such a block: then one launch
        produces both y (memoised for the shortcut / any other consumer) and that convolution's output."""
        nxt = None
        ysym = sym
        if self._plain_conv(e, 256, 64, 1, 0, "relu", False) or self._plain_conv(e, 256, 128, 1, 0, "relu", False):
            # 64: the next block of the stage; 128: the first block of the next stage (not together with a downsample
            # shortcut: the two 32 KiB filters do not fit side by side - a one-block stage, which ResNet does not have)
            pm = self._bottleneck64_match(e.x.expr)
            if pm is not None and not (e.weight.shape[0] == 128 and pm[1] is not None):
                nxt, ysym = e, e.x
        m = self._bottleneck64_match(ysym.expr)
        if m is None:
            return None
        c3 = ysym.expr
        c2, down = m
        _, h, wd = ysym.shape
        t1 = self.emit(c2.x)          # first: it may itself be the fused tail of the previous block (which memoises y)
        if t1.pitch != 64:
            return None
        w2, b2 = _pack.fold_bn(c2.weight, c2.bias, c2.bn)
        w3, b3 = _pack.fold_bn(c3.weight, c3.bias, c3.bn)
        w3p = _pack.pack_conv_weight(w3, 64)
        res = x0 = None
        if down is not None:
            x0 = self.emit(down.x)
            wdn, bdn = _pack.fold_bn(down.weight, down.bias, down.bn)
            w3p = torch.cat([w3p, _pack.pack_conv_weight(wdn, 64)], dim=1).contiguous()
            b3 = b3 + bdn
            if x0.pitch != 64:
                return None
        else:
            res = self.emit(c3.res)
            if res.pitch != 256:
                return None
        y = self.alloc(self.n * h * wd, 256, (h, wd))
        kw = dict(t1=t1.map(h, wd), w2=self.const(_pack.pack_conv_weight(w2, 64)), b2=self.const(b2), w3=self.const(w3p),
                  b3=self.const(b3), residual=None if res is None else res.map(h, wd),
                  x0=None if x0 is None else x0.map(h, wd), out=y.map(h, wd))
        out = y
        if nxt is not None:
            w1, b1 = _pack.fold_bn(nxt.weight, nxt.bias, nxt.bn)
            out = self.alloc(self.n * h * wd, nxt.weight.shape[0], (h, wd))
            kw.update(w1n=self.const(_pack.pack_conv_weight(w1, 256)), b1n=self.const(b1), next_out=out.map(h, wd))
            self.keep.append(c3)
            self.memo[id(c3)] = y
        self.keep.append(c2)
        self.step(ops.bottleneck64, **kw)
        return out

    def _emit_grouped(self, sym, e, w, b, act, res_after, res, dst, out_f32):
        """ResNeXt conv2 (resnet.py:83 with groups=32): block-diagonal implicit GEMM over 64-channel blocks
        (EQXV_FLAG_GROUPED_BLOCK64): C/64 independent dense 64->64 convolutions in one launch."""
        cout, cin_g, kh, kw = e.weight.shape
        c_in, h, wd = e.x.shape
        (sh, sw), (ph, pw), (dh, dw) = e.stride, e.padding, e.dilation
        if sh != sw or ph != pw or dh != dw or out_f32:
            raise NotImplementedError("anisotropic / fp32-output grouped convolution")
        xb = self.emit(e.x)
        ho, wo = sym.shape[1:]
        if cout != c_in or c_in % 64 != 0 or 64 % cin_g != 0:
            # group geometries outside the 64-channel block layout (RegNet: group widths 8/16/24/48... on stage
            # widths such as 104 or 440, regnet.py:58-81): the filter is expanded to its dense block-diagonal
            # [Cout, Cin] form on the host and runs as an ordinary implicit GEMM (zeros cost flops, not correctness)
            cin_eff = xb.cpad
            wp = self.const(_pack.pack_conv_weight(_pack.expand_grouped_weight(w, e.groups), cin_eff))
            out = dst if dst is not None else self.alloc(self.n * ho * wo, cout, (ho, wo))
            bias_d = self.const(b) if b is not None else None
            omap, rmap = out.map(ho, wo), None if res is None else res.map(ho, wo)
            for n0, n1 in _n_chunks(cout):
                self.step(ops.conv2d, x=xb.map(h, wd, cin_eff), wgt=wp[n0:n1],
                          bias=None if bias_d is None else bias_d[n0:n1], cin=cin_eff, cout=n1 - n0, kh=kh,
                          kw=kw, stride=sh, pad=ph, dil=dh, act=act,
                          residual=None if rmap is None else rmap[..., n0:n1], res_after_act=res_after,
                          out=omap if (n0, n1) == (0, cout) else omap[..., n0:n1])
            return out
        out = dst if dst is not None else self.alloc(self.n * ho * wo, cout, (ho, wo))
        wp = self.const(_pack.pack_grouped_weight(w, e.groups))
        bias_d = self.const(b) if b is not None else None
        self.step(ops.conv2d, x=xb.map(h, wd, c_in), wgt=wp, bias=bias_d, cin=c_in, cout=cout, kh=kh, kw=kw,
                  stride=sh, pad=ph, dil=dh, act=act, residual=None if res is None else res.map(ho, wo),
                  res_after_act=res_after, out=out.map(ho, wo), grouped_block64=True)
        return out

    def _emit_depthwise(self, sym, e, w, b, act, res, dst):
        """groups == channels: shared-memory-free channels-last stencil (csrc/depthwise.cu)"""
        if res is not None:
            raise NotImplementedError("residual on a depthwise convolution")
        c, h, wd = e.x.shape
        k = e.weight.shape[2]
        (sh, sw), (ph, pw), (dh, dw) = e.stride, e.padding, e.dilation
        if e.weight.shape[2] != e.weight.shape[3] or sh != sw or ph != pw or dh != dw:
            raise NotImplementedError("anisotropic depthwise convolution")
        xb = self.emit(e.x)
        cp = xb.cpad
        wp = self.const(_pack.pack_depthwise_weight(w, cp))
        bias = torch.zeros(cp)
        if b is not None:
            bias[:c] = b
        bias_d = self.const(bias)
        ho, wo = sym.shape[1:]
        out = dst if dst is not None else self.alloc(self.n * ho * wo, c, (ho, wo))
        if id(e) in self._pool_requests and self._dw_pool_fusable(e, act):
            # SqueezeExcitation right behind the depthwise conv (efficientnet.py:138-160): the squeeze (squeeze.py:52)
            # comes out of the same kernel, no pass over the expanded tensor just to average it
            pooled = self._pool_requests.pop(id(e))
            need = ops.dwconv_pool_workspace_bytes(self.n, h, wd, cp, k, sh, ph) if self.device.type == "cuda" else 0
            ws = torch.zeros(max(need, 16), dtype=torch.uint8, device=self.device) if need else None
            self.step(ops.dwconv_pool, x=xb.map(h, wd, cp), wgt=wp, bias=bias_d, k=k, stride=sh, pad=ph, act=act,
                      out=out.map(ho, wo, cp), pooled=pooled.rows(cp), workspace=ws)
            return out
        self.step(ops.dwconv, x=xb.map(h, wd, cp), wgt=wp, bias=bias_d, k=k, stride=sh, pad=ph, dil=dh, act=act,
                  out=out.map(ho, wo, cp))
        return out

    @staticmethod
    def _dw_pool_fusable(e: "T.Conv", act: Optional[int] = None) -> bool:
        k = e.weight.shape[2]
        ok = (e.groups == e.weight.shape[0] and e.weight.shape[1] == 1 and e.groups > 1 and e.res is None
              and k == e.weight.shape[3] and k in (3, 5) and e.stride in ((1, 1), (2, 2)) and e.dilation == (1, 1)
              and e.padding[0] == e.padding[1] and e.weight.shape[0] % 8 == 0)
        if act is not None:
            ok = ok and act in (_lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_SILU, _lib.ACT_HARDSWISH)
        return ok

    # ---- linear ------------------------------------------------------------------------------
    # The LayerNorm between two GEMMs (vit.py:149,154) is folded into them: the producer's epilogue emits row statistics,
    # the consumer applies mean / rstd to its accumulators (include/eqxv_b200.h, K5 + K7). EQXV_NO_LN_FOLD=1: A/B switch.
    LN_FOLD = os.environ.get("EQXV_NO_LN_FOLD") != "1"
    GATE_FUSE = os.environ.get("EQXV_NO_GATE_FUSE") != "1"   # SE gate inside the projection GEMM (A/B switch)
    LN_MAX_COUT = 3328    # the consumer stages bias AND filter column sums (28 KB area: csrc/igemm.cu)

    def _emit_linear_ln(self, sym, e: T.Linear, ln: "T.LayerNormE", act: int) -> Optional[Buf]:
        """Linear(LayerNorm(x)) with x produced by a residual GEMM of this plan: no LayerNorm kernel, no LayerNorm
        output in HBM. Returns None when the pattern does not apply (the caller lowers the two ops separately)."""
        out_f, in_f = e.weight.shape
        pkey = id(ln.x.expr)
        xb = self.emit(ln.x)
        idx = self._ln_producers.get(pkey)
        if idx is None or xb.pitch != in_f or in_f % 8 != 0:
            return None
        rows = xb.t.shape[0]
        stats = self._ln_stats.get(pkey)
        if stats is None:
            stats = self.arena_ok(torch.zeros((rows, (in_f + 63) // 64, 2), dtype=torch.float32, device=self.device))
            self.act_bytes += stats.numel() * 4
            self._ln_stats[pkey] = stats
            fn, kw = self.steps[idx]
            assert fn is ops.gemm
            self.steps[idx] = (ops.gemm_rowstats, dict(kw, stats=stats))
        gamma, beta = ln.weight.detach().double(), ln.bias.detach().double()
        w64 = e.weight.detach().double()
        wp = _pack.pack_linear_weight((w64 * gamma[None, :]).float(), in_f)       # [out_f, in_f] bf16, gamma folded
        wsum = wp.float().sum(1)                                                  # of the ROUNDED filter the MMA sees
        bias = (w64 @ beta + (e.bias.detach().double().reshape(-1) if e.bias is not None else 0.0)).float()
        wp_d, wsum_d, bias_d = self.const(wp), self.const(wsum), self.const(bias)
        geom = (sym.shape[0],) if sym.kind == "tokens" else ()
        out = self.alloc(rows, out_f, geom)
        o2d = out.rows()
        for n0, n1 in _n_chunks(out_f, self.LN_MAX_COUT):
            self.step(ops.gemm_ln, a=xb.rows(in_f), wgt=wp_d[n0:n1], bias=bias_d[n0:n1], wsum=wsum_d[n0:n1], stats=stats,
                      eps=float(ln.eps), act=act, out=o2d if (n0, n1) == (0, out_f) else o2d[:, n0:n1])
        return out

    def _emit_Linear(self, sym, e: T.Linear, out_f32=False):
        out_f, in_f = e.weight.shape
        act, res_after = self._epilogue(e)
        w = e.weight.detach().float()
        xsym = e.x
        if (self.LN_FOLD and isinstance(xsym.expr, T.LayerNormE) and not out_f32 and e.res is None
                and act in (_lib.ACT_NONE, _lib.ACT_GELU_TANH) and id(xsym.expr) not in self.memo):
            folded = self._emit_linear_ln(sym, e, xsym.expr, act)
            if folded is not None:
                return folded
        if isinstance(xsym.expr, T.Ravel) and xsym.expr.x.shape[1] * xsym.expr.x.shape[2] > 1:
            # flatten of a (C,H,W) map is in C,H,W order (jnp.ravel, vgg.py:116); the buffer is H,W,C
            src = xsym.expr.x
            c, h, wd = src.shape
            xb = self.emit(src)
            if xb.pitch != c:
                raise NotImplementedError("flatten of a channel-padded map")
            w = w.reshape(out_f, c, h, wd).permute(0, 2, 3, 1).reshape(out_f, in_f)
            a = torch.as_strided(xb.t, (self.n, in_f), (in_f, 1), xb.t.storage_offset())
            k = in_f
            rows = self.n
        else:
            xb = self.emit(xsym)
            k = xb.cpad
            a = xb.rows(k)
            rows = a.shape[0]
        wp = self.const(_pack.pack_linear_weight(w, k))
        bias_d = self.const(e.bias.detach().float().reshape(-1)) if e.bias is not None else None
        res = self.emit(e.res) if e.res is not None else None
        geom = (sym.shape[0],) if sym.kind == "tokens" else ()
        out = self.alloc(rows, out_f, geom, dtype=torch.float32 if out_f32 else BF16)
        o2d, r2d = out.rows(out_f if out_f32 else None), None if res is None else res.rows()
        chunks = _n_chunks(out_f)
        for n0, n1 in chunks:
            self.step(ops.gemm, a=a, wgt=wp[n0:n1], bias=None if bias_d is None else bias_d[n0:n1], act=act,
                      residual=None if r2d is None else r2d[:, n0:n1], res_after_act=res_after,
                      out=o2d[:, n0:n1] if (n0, n1) != (0, out_f) else o2d, out_f32=out_f32)
        if (len(chunks) == 1 and res is not None and act == _lib.ACT_NONE and not res_after and not out_f32
                and bias_d is not None and out.pitch == out_f):
            self._ln_producers[id(sym.expr)] = len(self.steps) - 1   # may later emit row statistics for a LayerNorm
        return out

    # ---- shape-only nodes --------------------------------------------------------------------
    def _emit_Ravel(self, sym, e):
        src = self.emit(e.x)
        c, h, w = e.x.shape
        if h * w != 1:
            raise NotImplementedError("ravel of a spatial map is only supported in front of a Linear")
        return Buf(src.t, src.c, self.n, ())

    def _emit_ToTokens(self, sym, e):
        x = e.x
        ce = x.expr
        if isinstance(ce, T.Conv) and isinstance(ce.x.expr, T.Input) and ce.groups == 1 \
                and ce.stride == tuple(ce.weight.shape[2:]) and ce.padding == (0, 0) and ce.dilation == (1, 1) \
                and ce.weight.shape[2] == ce.weight.shape[3] and ce.weight.shape[2] % 8 == 0 \
                and ce.res is None:
            return self._emit_patch_embed(sym, ce)
        src = self.emit(x)
        c, h, w = x.shape
        return Buf(src.t, src.c, self.n, (h * w,))

    def _emit_patch_embed(self, sym, ce: T.Conv):
        """PatchEmbed (patch_embed.py:79-82): conv PxP stride P == GEMM over patch rows"""
        d, cin, p, _ = ce.weight.shape
        c, h, w = self.in_shape
        np_ = (h // p) * (w // p)
        w_f, b_f = _pack.fold_bn(ce.weight, ce.bias, ce.bn)
        act, _ = self._epilogue(ce)
        k = cin * p * p
        rows = self.arena_ok(torch.empty((self.n * np_, k), dtype=BF16, device=self.device))
        self.act_bytes += rows.numel() * 2
        if self.u8 is not None:
            self.step(ops.u8_patchify, x=self.x_in, lut=self.lut, p=p, out=rows)
        else:
            self.step(ops.patchify, x_nchw=self.x_in, p=p, out=rows)
        wp = self.const(w_f.reshape(d, k).to(torch.bfloat16))
        bias_d = self.const(b_f) if b_f is not None else None
        out = self.alloc(self.n * np_, d, (np_,))
        self.step(ops.gemm, a=rows, wgt=wp, bias=bias_d, act=act, out=out.rows())
        return out

    def _emit_ToMap(self, sym, e):
        src = self.emit(e.x)
        return Buf(src.t, src.c, self.n, (e.h, e.w))

    # ---- pooling -----------------------------------------------------------------------------
    # Measured on B200 (tools/bench_tail.py, batch 256): stem 174 us + max-pool 118 us separately, 401 us fused - the 192
    # red.global.max (bf16 x 8) per tile that combine the pooled pixels on tile borders cost more than the 822 MB of
    # traffic the fusion removes. The kernel stays (bit-exact, tested) but the lowering uses it only on request.
    STEM_POOL_FUSE = os.environ.get("EQXV_STEM_POOL") == "1"

    def _emit_stem_maxpool(self, sym, e: T.Pool) -> Optional[Buf]:
        """conv1 -> bn1 -> relu -> maxpool 3x3 / 2 / 1 (resnet.py:243-253) in ONE kernel: the convolution's output is
        never materialised. Applies to a first-layer convolution with 64 filters whose output tiles into 16 x 8 blocks."""
        ce = e.x.expr
        if not (self.STEM_POOL_FUSE and e.mode == "max" and (e.k, e.stride, e.pad) == (3, 2, 1) and not e.ceil):
            return None
        if not (isinstance(ce, T.Conv) and id(ce) not in self.memo and isinstance(ce.x.expr, T.Input) and ce.groups == 1):
            return None
        cout, cin, kh, kw = ce.weight.shape
        (sh, sw), (ph, pw), (dh, dw) = ce.stride, ce.padding, ce.dilation
        _, hc, wc = e.x.shape
        if not (cout == 64 and cin <= 8 and kh <= 8 and kw <= 8 and sh == sw and sh in (1, 2, 4) and ph == pw and dh == dw == 1
                and 2 * ph <= kw and ce.res is None and ce.act1 == "relu" and ce.act2 is None and hc % 16 == 0 and wc % 8 == 0):
            return None
        w, b = _pack.fold_bn(ce.weight, ce.bias, ce.bn)
        if b is None:
            return None
        _, h, wd = ce.x.shape
        ho, wo = sym.shape[1:]
        out = self.alloc(self.n * ho * wo, cout, (ho, wo))
        self.step(ops.conv_stem_maxpool, xpad=self._stem_input(ph, h, wd), wgt=self.const(_pack.pack_stem_weight(w)),
                  bias=self.const(b), n=self.n, h=h, w=wd, cout=cout, kh=kh, kw=kw, stride=sh, pad=ph, out=out.map(ho, wo))
        return out

    def _emit_Pool(self, sym, e: T.Pool, dst: Optional[Buf] = None):
        if dst is None:
            fused = self._emit_stem_maxpool(sym, e)
            if fused is not None:
                return fused
        xb = self.emit(e.x)
        c, h, w = e.x.shape
        ho, wo = sym.shape[1:]
        out = dst if dst is not None else self.alloc(self.n * ho * wo, c, (ho, wo))
        if e.mode == "max":
            self.step(ops.maxpool2d, x=xb.map(h, w), k=e.k, stride=e.stride, pad=e.pad, out=out.map(ho, wo),
                      ceil_mode=e.ceil)
        else:
            self.step(ops.avgpool2d, x=xb.map(h, w), k=e.k, stride=e.stride, out=out.map(ho, wo))
        return out

    def _emit_AdaptiveAvgPool(self, sym, e: T.AdaptiveAvgPool):
        c, h, w = e.x.shape
        xe = e.x.expr
        if (e.oh, e.ow) == (1, 1) and isinstance(xe, T.Conv) and id(xe) not in self.memo and self._dw_pool_fusable(xe):
            pooled = self.alloc(self.n, c, (1, 1))
            self._pool_requests[id(xe)] = pooled  # (emission recurses: several requests can be open at once)
            self.emit(e.x)
            if self._pool_requests.pop(id(xe), None) is None:   # taken: the depthwise kernel writes the squeeze as well
                return pooled
            self.step(ops.adaptive_avgpool, x=self.memo[id(xe)].map(h, w), oh=1, ow=1, out=pooled.map(1, 1))
            return pooled
        xb = self.emit(e.x)
        out = self.alloc(self.n * e.oh * e.ow, c, (e.oh, e.ow))
        self.step(ops.adaptive_avgpool, x=xb.map(h, w), oh=e.oh, ow=e.ow, out=out.map(e.oh, e.ow))
        return out

    # ---- transformer pieces ------------------------------------------------------------------
    def _emit_LayerNormE(self, sym, e: T.LayerNormE):
        xb = self.emit(e.x)
        d = sym.shape[-1]
        if d % 8 != 0:
            raise NotImplementedError("LayerNorm width must be a multiple of 8")
        out = self.alloc(xb.t.shape[0], d, xb.geom)
        g = self.const(e.weight.detach().float())
        b = self.const(e.bias.detach().float())
        self.step(ops.layernorm, x=xb.rows(d), gamma=g, beta=b, eps=float(e.eps), out=out.rows(d))
        return out

    def _emit_Attention(self, sym, e: T.Attention):
        qkv = self.emit(e.qkv)
        t, c = sym.shape
        hd = c // e.heads
        if qkv.pitch != 3 * c:
            raise NotImplementedError("attention expects a dense qkv matrix")
        out = self.alloc(self.n * t, c, (t,))
        self.step(ops.attention, qkv=qkv.rows(3 * c), images=self.n, tokens=t, heads=e.heads, head_dim=hd,
                  scale=float(e.scale), out=out.rows(c))
        return out

    def _emit_WindowAttention(self, sym, e: T.WindowAttention):
        """Swin attention core (swin.py:117-253): roll / window partition / mask are index arithmetic
        inside eqxv_window_attention_bf16, nothing is permuted in memory."""
        qkv = self.emit(e.qkv)
        t, c = sym.shape
        hd = c // e.heads
        if qkv.pitch != 3 * c:
            raise NotImplementedError("window attention expects a dense qkv matrix")
        if e.window[0] != e.window[1]:
            raise NotImplementedError("non-square attention windows")
        out = self.alloc(self.n * t, c, (t,))
        bias = self.const(e.bias.detach().float().contiguous())
        if e.cosine_scale is not None:
            # Swin-V2 (swin.py:158-166): normalise q, k over the windows of each image, scale q per head - in place on
            # the qkv matrix (its only consumer is this attention)
            sq = self.const(e.cosine_scale.detach().float().reshape(-1).contiguous())
            self.step(ops.swin_v2_qk_normalize, qkv=qkv.rows(3 * c), scale_q=sq, n=self.n, h=e.h, w=e.w,
                      heads=e.heads, head_dim=hd, window=e.window[0], shift=e.shift)
        self.step(ops.window_attention, qkv=qkv.rows(3 * c), bias=bias, n=self.n, h=e.h, w=e.w, heads=e.heads,
                  head_dim=hd, window=e.window[0], shift=e.shift, scale=float(e.scale), out=out.rows(c))
        return out

    def _emit_PatchMerge(self, sym, e: T.PatchMerge):
        xb = self.emit(e.x)
        c, h, w = e.x.shape
        if c % 8 != 0:
            raise NotImplementedError("patch merging needs a channel count that is a multiple of 8")
        out = self.alloc(self.n * (h // 2) * (w // 2), 4 * c, (h // 2, w // 2))
        self.step(ops.patch_merge, x=xb.map(h, w), out=out.map(h // 2, w // 2))
        return out

    def _emit_AttentionProbs(self, sym, e):
        raise EqxvError("the attention matrix (vit.py:151-152) can only be a model output")

    def _emit_ClsPos(self, sym, e: T.ClsPos):
        xb = self.emit(e.x)
        t, d = sym.shape
        out = self.alloc(self.n * t, d, (t,))
        cls = self.const(e.cls.detach().float().reshape(-1))
        pos = self.const(e.pos.detach().float().reshape(t, d))
        self.step(ops.vit_assemble_tokens, patches=xb.rows(d), cls=cls, pos=pos, n=self.n, np_=t - 1, d=d,
                  out=out.rows(d))
        return out

    def _emit_SelectRow(self, sym, e: T.SelectRow):
        xb = self.emit(e.x)
        t, d = e.x.shape
        out = self.alloc(self.n, d, ())
        self.step(ops.gather_rows, x=xb.rows(d), n=self.n, tokens=t, row=e.row, out=out.rows(d))
        return out

    # ---- elementwise nodes that could not be folded into a GEMM epilogue ----------------------
    def _padded(self, v: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.zeros(n, dtype=torch.float32)
        out[: v.numel()] = v.detach().float().reshape(-1)
        return self.const(out)

    def _like(self, src: Buf, sym) -> Buf:
        return self.alloc(src.t.shape[0], sym.shape[0] if sym.kind == "chw" else sym.shape[-1], src.geom)

    def _emit_BNAct(self, sym, e: T.BNAct):
        """standalone inference BatchNorm (+activation), densenet.py:64-65,118,211"""
        xb = self.emit(e.x)
        scale, shift = e.bn.folded()
        out = self._like(xb, sym)
        self.step(ops.eltwise, x=xb.rows(), scale=self._padded(scale, xb.cpad), shift=self._padded(shift, xb.cpad),
                  act=ACT_BY_NAME[e.act], out=out.rows())
        return out

    def _emit_Act(self, sym, e: T.Act):
        xb = self.emit(e.x)
        out = self._like(xb, sym)
        self.step(ops.eltwise, x=xb.rows(), act=ACT_BY_NAME[e.act], out=out.rows())
        return out

    def _emit_Add(self, sym, e: T.Add):
        a, b = self.emit(e.a), self.emit(e.b)
        out = self._like(a, sym)
        self.step(ops.eltwise, x=a.rows(), other=b.rows(), act=ACT_BY_NAME[e.act], out=out.rows())
        return out

    def _emit_ChannelScale(self, sym, e: T.ChannelScale):
        """x * s with one gate per (image, channel): the SqueezeExcitation output (squeeze.py:61)"""
        sb = self.emit(e.s)      # the gate first: its squeeze may ride in the kernel that produces x (depthwise + pool)
        xb = self.emit(e.x)
        c, h, w = e.x.shape
        out = self._like(xb, sym)
        self.step(ops.eltwise, x=xb.rows(), gate=sb.rows(xb.cpad), rows_per_image=h * w, out=out.rows())
        return out

    def _emit_Concat(self, sym, e: T.Concat):
        """Channel concatenation without a concat pass: producers store into channel slices of one
        buffer. Concats that extend an earlier concat (DenseNet: [x0,f1] -> [x0,f1,f2] -> ...) keep
        growing inside the same buffer, sized by the `capacity` hint of the model code."""
        c_tot, h, w = sym.shape
        rows = self.n * h * w
        gkey = id(e.xs[0].expr)
        grp = self._concat_groups.get(gkey)
        ids = [id(x.expr) for x in e.xs]
        if grp is None or grp["cap"] < c_tot or grp["ids"] != ids[: len(grp["ids"])]:
            cap = max(c_tot, e.capacity or 0)
            base = torch.zeros((rows, _round8(cap)), dtype=BF16, device=self.device)
            self.act_bytes += base.numel() * 2
            grp = {"base": base, "cap": cap, "ids": [], "width": 0}
            self._concat_groups[gkey] = grp
        base = grp["base"]
        for x in e.xs[len(grp["ids"]):]:
            cx = x.shape[0]
            off = (grp["width"] + e.align - 1) // e.align * e.align
            if off % 8 != 0:
                raise NotImplementedError("concat slices must start at a multiple of 8 channels")
            sl = Buf(torch.as_strided(base, (rows, cx), (base.stride(0), 1), base.storage_offset() + off), cx,
                     self.n, (h, w))
            self.emit(x, dst=sl)
            grp["ids"].append(id(x.expr))
            grp["width"] = off + cx
        view = torch.as_strided(base, (rows, c_tot), (base.stride(0), 1), base.storage_offset())
        return Buf(view, c_tot, self.n, (h, w))

    def _emit_ChannelView(self, sym, e: T.ChannelView, dst: Optional[Buf] = None):
        """A channel gather that has to exist in memory (the pass-through half of a ShuffleNetV2 unit,
        shufflenetv2.py:134-135, or the input of a depthwise conv): a 1x1 convolution with a 0/1 selection matrix on
        the tensor-core GEMM. Exact: each output is one bf16 input times 1.0 in an fp32 accumulator. Dense
        convolutions never get here, they absorb the gather into their filter (`_trace.conv2d`)."""
        sel = torch.zeros((len(e.idx), e.x.shape[0], 1, 1), dtype=torch.float32)
        sel[torch.arange(len(e.idx)), torch.tensor(e.idx, dtype=torch.long)] = 1.0
        conv = T.Conv(e.x, sel, None, (1, 1), (0, 0), (1, 1), 1)
        self.keep.append(conv)
        return self._emit_Conv(sym, conv, False, dst)

    def _emit_Resize(self, sym, e: T.Resize, dst: Optional[Buf] = None):
        """bilinear upsample kept in NHWC bf16 (ASPP pooling branch broadcast, deeplabv3.py:74)"""
        xb = self.emit(e.x)
        c, h, w = e.x.shape
        out = dst if dst is not None else self.alloc(self.n * e.h * e.w, c, (e.h, e.w))
        self.step(ops.resize_bilinear, x=xb.map(h, w), oh=e.h, ow=e.w, out=out.map(e.h, e.w, xb.cpad))
        return out

    # -------------------------------------------------------------- outputs
    def add_output(self, sym: T.Sym):
        """Materialise `sym` as an fp32 tensor in the reference's per-sample layout, batched."""
        n = self.n
        if sym.kind == "vec" and isinstance(sym.expr, T.Linear) and id(sym.expr) not in self.memo:
            buf = self.emit(sym, out_f32=True)
            self.outputs.append((buf.rows(sym.shape[0]), (n,) + sym.shape))
            return
        if sym.kind == "chw" and isinstance(sym.expr, T.Resize) and id(sym.expr) not in self.memo:
            # segmentation output (_utils.py:52,57): upsample straight into the fp32 NCHW result
            src = self.emit(sym.expr.x)
            c, hs, ws = sym.expr.x.shape
            out = torch.empty((n, c, sym.expr.h, sym.expr.w), dtype=torch.float32, device=self.device)
            self.step(ops.resize_bilinear_to_nchw, x=src.map(hs, ws), c=c, oh=sym.expr.h, ow=sym.expr.w, out=out)
            self.outputs.append((out, (n,) + sym.shape))
            return
        if sym.kind == "attn" and isinstance(sym.expr, T.AttentionProbs):
            # _VitBlock(return_attention=True) / get_last_self_attention (vit.py:151-152, 275-292):
            # fp32 probabilities (1, heads, T, T) per sample, written directly by the kernel
            e = sym.expr
            qkv = self.emit(e.qkv)
            _, heads, t, _ = sym.shape
            c = qkv.pitch // 3
            if qkv.pitch != 3 * heads * (c // heads):
                raise NotImplementedError("attention expects a dense qkv matrix")
            out = torch.empty((n, 1, heads, t, t), dtype=torch.float32, device=self.device)
            self.step(ops.attention_probs, qkv=qkv.rows(3 * c), images=n, tokens=t, heads=heads,
                      head_dim=c // heads, scale=float(e.scale), out=out)
            self.outputs.append((out, (n,) + sym.shape))
            return
        buf = self.emit(sym)
        if sym.kind == "chw":
            c, h, w = sym.shape
            out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
            self.step(ops.nhwc_to_nchw, x=buf.map(h, w), c=c, out=out)
        elif sym.kind in ("vec", "tokens"):
            d = sym.shape[-1]
            rows = buf.t.shape[0]
            out = torch.empty((rows, d), dtype=torch.float32, device=self.device)
            self.step(ops.nhwc_to_nchw, x=buf.rows(d).view(rows, 1, 1, d) if buf.pitch == d else
                      torch.as_strided(buf.t, (rows, 1, 1, d), (buf.pitch, buf.pitch, buf.pitch, 1),
                                       buf.t.storage_offset()), c=d, out=out.view(rows, d, 1, 1))
            out = out.view((n,) + sym.shape)
        else:
            raise NotImplementedError(f"output of kind {sym.kind}")
        self.outputs.append((out, (n,) + sym.shape))

    # -------------------------------------------------------------- execution
    def run_steps(self, stream: int):
        for fn, kw in self.steps:
            fn(stream=stream, **kw)

    def capture(self, stream: int):
        _lib.call("eqxv_graph_begin", stream)
        g = C.c_void_p()
        try:
            self.run_steps(stream)
        except BaseException:
            # leave capture mode, drop whatever was recorded, and surface the ORIGINAL error
            try:
                _lib.call("eqxv_graph_end", stream, C.byref(g))
                if g.value:
                    _lib.call("eqxv_graph_destroy", g)
            except EqxvError:
                pass
            raise
        _lib.call("eqxv_graph_end", stream, C.byref(g))
        self.graph = g

    def launch(self, stream: int):
        if self.graph is not None:
            _lib.call("eqxv_graph_launch", self.graph, stream)
        else:
            self.run_steps(stream)

    @property
    def num_launches(self) -> int:
        return len(self.steps)

    # -------------------------------------------------------------- lanes
    def _owned_storages(self) -> Dict[int, torch.Tensor]:
        """every tensor of the plan that is NOT a constant (activations, input, outputs), by storage address"""
        const_ptrs = {c.untyped_storage().data_ptr() for c in self.consts}
        found: Dict[int, torch.Tensor] = {}

        def visit(t):
            if isinstance(t, torch.Tensor) and t.device.type != "meta":
                p_ = t.untyped_storage().data_ptr()
                if p_ not in const_ptrs and p_ not in found:
                    found[p_] = t

        for _, kw in self.steps:
            for v in kw.values():
                visit(v)
        for t in (self.x_in, self.x_host_target):
            visit(t)
        for o, _ in self.outputs:
            visit(o)
        return found

    def clone_lane(self) -> "Plan":
        """A second instance of this plan on its own buffers (same constants, same steps): used round-robin with the
        original so that the host-to-device copy of call i+1 overlaps the kernels of call i."""
        owned = self._owned_storages()
        fresh: Dict[int, torch.UntypedStorage] = {}
        for p_, t in owned.items():
            st = t.untyped_storage()
            nb = torch.empty(st.nbytes(), dtype=torch.uint8, device=t.device)
            nb.copy_(torch.empty(0, dtype=torch.uint8, device=t.device).set_(st, 0, (st.nbytes(),), (1,)))
            fresh[p_] = nb.untyped_storage()   # keeps the zero-filled pad columns / concat slots of the original

        def remap(t):
            if not isinstance(t, torch.Tensor):
                return t
            st = fresh.get(t.untyped_storage().data_ptr())
            if st is None:
                return t
            return torch.empty(0, dtype=t.dtype, device=t.device).set_(st, t.storage_offset(), t.shape, t.stride())

        lane = Plan.__new__(Plan)
        lane.__dict__.update({k: v for k, v in self.__dict__.items()
                              if k not in ("steps", "outputs", "graph", "stream", "done", "out_ready", "x_in",
                                           "x_host_target")})
        lane.steps = [(fn, {k: remap(v) for k, v in kw.items()}) for fn, kw in self.steps]
        lane.outputs = [(remap(o), shp) for o, shp in self.outputs]
        lane.x_in, lane.x_host_target = remap(self.x_in), remap(self.x_host_target)
        lane.graph = lane.stream = lane.done = lane.out_ready = None
        return lane

    def close(self):
        """release the CUDA objects of the plan (graph exec, lane stream, events); buffers die with the object"""
        if self.stream is not None:
            try:
                _lib.call("eqxv_stream_sync", self.stream)
            except EqxvError:
                pass
        if self.graph is not None:
            _lib.call("eqxv_graph_destroy", self.graph)
            self.graph = None
        if getattr(_state, "last_done", None) is self.done:
            _state.last_done = None        # the FIFO chain must not wait on a destroyed event
        for name in ("done", "out_ready"):
            ev = getattr(self, name, None)
            if ev is not None:
                _lib.call("eqxv_event_destroy", ev)
                setattr(self, name, None)
        if self.stream is not None:
            _lib.call("eqxv_stream_destroy", self.stream)
            self.stream = None


# ------------------------------------------------------------------------------------------------
# engine: plan cache, streams, public entry points
# ------------------------------------------------------------------------------------------------
_state = threading.local()


def _ctx():
    if not hasattr(_state, "stream"):
        if not torch.cuda.is_available():
            raise EqxvError("eqxvision_b200 needs a CUDA device (sm_100a): there is no CPU fallback")
        dev = torch.cuda.current_device()
        _lib.init(dev)
        _state.stream = _new_stream()
        _state.device = torch.device("cuda", dev)
        _state.last_done = None      # event of the most recent graph launch of this thread (FIFO chaining)
        _state.fence = _new_event()  # orders a lane stream after torch's current stream
    return _state


def _new_stream() -> int:
    s = C.c_void_p()
    _lib.call("eqxv_stream_create", C.byref(s))
    return s.value


def _new_event():
    e = C.c_void_p()
    _lib.call("eqxv_event_create", C.byref(e))
    return e


def stream_handle() -> int:
    return _ctx().stream


def _flatten_out(out, acc: List[T.Sym]):
    if T.is_sym(out):
        acc.append(out)
        return ("leaf", len(acc) - 1)
    if out is None:   # `(None, out)` of segmentation models without an auxiliary head (_utils.py:54, lraspp.py:68)
        return ("none", None)
    if isinstance(out, (list, tuple)):
        return (type(out).__name__, [_flatten_out(o, acc) for o in out])
    raise EqxvError(f"model returned an unsupported value: {out!r}")


def _unflatten_out(struct, leaves):
    tag, payload = struct
    if tag == "none":
        return None
    if tag == "leaf":
        return leaves[payload]
    items = [_unflatten_out(s, leaves) for s in payload]
    return tuple(items) if tag == "tuple" else items


def _arm_lane(plan: Plan, stream: Optional[int], use_graph: bool) -> Plan:
    """warm-up run (also surfaces launch errors outside capture), then capture, on the lane's own stream"""
    plan.stream = _new_stream() if stream is None else stream
    plan.done, plan.out_ready = _new_event(), _new_event()
    # buffers were zero-filled / constants uploaded on torch's current stream: order them before the first launch
    torch.cuda.current_stream().synchronize()
    plan.run_steps(plan.stream)
    _lib.call("eqxv_stream_sync", plan.stream)
    if use_graph:
        plan.capture(plan.stream)
    return plan


def build_plan(module, method: str, batch: int, in_shape: Tuple[int, ...], args=(), kwargs=None,
               use_graph: bool = True, u8: Optional[dict] = None, stream: Optional[int] = None) -> Plan:
    ctx = _ctx()
    kwargs = dict(kwargs or {})
    # the batched call receives one key per sample (keys[B,2] in the reference); the traced
    # per-sample function only needs "a key or None" (keys are dead in inference)
    kwargs["key"] = None if kwargs.get("key") is None else jrandom.PRNGKey(0)
    plan = Plan(ctx.device, batch, in_shape, u8=u8)
    kind = {3: "chw", 2: "tokens", 1: "vec"}.get(len(in_shape))
    if kind is None:
        raise EqxvError(f"unsupported per-sample input rank {len(in_shape)}")
    x = T.Sym(kind, in_shape, T.Input())
    if kind != "chw":
        plan.memo[id(x.expr)] = _token_input(plan, kind, in_shape)
    fn = getattr(type(module), method)
    fn = getattr(fn, "__wrapped__", fn)
    out = fn(module, x, *args, **kwargs)
    syms: List[T.Sym] = []
    plan.out_struct = _flatten_out(out, syms)
    for s in syms:
        plan.add_output(s)
    plan.memo.clear()
    plan.keep.clear()
    plan.plan_memory()
    return _arm_lane(plan, ctx.stream if stream is None else stream, use_graph)


def _token_input(plan: Plan, kind: str, in_shape) -> Buf:
    """token / vector inputs (sub-module tests such as _VitBlock on (T,D)): fp32 -> bf16 rows"""
    d = in_shape[-1]
    rows = plan.n * (in_shape[0] if kind == "tokens" else 1)
    if d % 8 != 0:
        raise NotImplementedError("token inputs need a feature width that is a multiple of 8")
    buf = plan.alloc(rows, d, (in_shape[0],) if kind == "tokens" else ())
    # [rows, d] fp32 "NCHW" with c=d, h=w=1 -> bf16 rows
    plan.step(ops.nchw_to_nhwc, x=plan.x_in.view(rows, d, 1, 1), c_pad=d, out=buf.rows(d).view(rows, 1, 1, d))
    return buf


class PlanSet:
    """the lanes of one cache entry: lane 0 is built by tracing, further lanes are clones on their own buffers and
    streams, created on demand when host-resident inputs make copy/compute overlap worthwhile"""

    def __init__(self, first: Plan):
        self.lanes: List[Plan] = [first]
        self.next = 0
        self.lock = threading.Lock()   # one enqueue at a time per entry: lanes share nothing else

    def pick(self, want_lanes: int) -> Plan:
        if self.next >= len(self.lanes) and len(self.lanes) < want_lanes:
            self.lanes.append(_arm_lane(self.lanes[0].clone_lane(), None, self.lanes[0].graph is not None))
        lane = self.lanes[self.next % len(self.lanes)]
        self.next = (self.next + 1) % max(want_lanes, 1)
        return lane

    def close(self):
        for p in self.lanes:
            p.close()
        self.lanes = []


MAX_PLANS_PER_MODULE = int(os.environ.get("EQXV_PLAN_CACHE", "8"))
PIPELINE_LANES = int(os.environ.get("EQXV_LANES", "3"))
BLOCKING = os.environ.get("EQXV_BLOCKING", "0") == "1"


def _plan_cache(module) -> "collections.OrderedDict":
    d = module.__dict__.get("_eqxv_plans")
    if d is None:
        d = collections.OrderedDict()
        object.__setattr__(module, "_eqxv_plans", d)
    return d


def _static_key(v):
    if isinstance(v, (bool, int, float, str, type(None))):
        return v
    return repr(type(v))


_cache_lock = threading.Lock()


def get_plan_set(module, method: str, batch: int, in_shape, args=(), kwargs=None, u8: Optional[dict] = None) -> PlanSet:
    kwargs = kwargs or {}
    ctx = _ctx()
    key = (method, batch, tuple(in_shape), tuple(_static_key(a) for a in args),
           tuple(sorted((k, _static_key(v)) for k, v in kwargs.items() if k != "key")),
           kwargs.get("key") is None,   # ResNet raises without a key (resnet.py:341-342): trace both ways
           None if u8 is None else (tuple(u8["mean"]), tuple(u8["std"]), tuple(u8.get("raw_hw") or ())),
           ctx.device.index)
    with _cache_lock:
        cache = _plan_cache(module)
        ps = cache.get(key)
        if ps is None:
            ps = PlanSet(build_plan(module, method, batch, tuple(in_shape), args, kwargs, u8=u8, stream=_new_stream()))
            cache[key] = ps
            while len(cache) > MAX_PLANS_PER_MODULE:     # LRU: serving many batch sizes must not leak HBM
                _, old = cache.popitem(last=False)
                old.close()
        else:
            cache.move_to_end(key)
        return ps


def get_plan(module, method: str, batch: int, in_shape, args=(), kwargs=None, u8: Optional[dict] = None) -> Plan:
    """lane 0 of the cache entry (benchmarks and tools drive it directly)"""
    return get_plan_set(module, method, batch, in_shape, args, kwargs, u8).lanes[0]


def clear_plans(module) -> None:
    """drop every cached plan of `module` (graphs, streams and activation buffers)"""
    with _cache_lock:
        cache = _plan_cache(module)
        while cache:
            _, ps = cache.popitem()
            ps.close()


def _enqueue_input(plan: Plan, x, ctx) -> None:
    """copy the caller's batch into the lane's input buffer on the lane's stream (no host synchronisation for pinned
    or device-resident sources; pageable sources are staged by the driver before the call returns)"""
    tgt = plan.x_host_target
    if plan.u8 is not None:
        t = x.pixels
    elif isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if t.dtype != tgt.dtype:
        t = t.to(tgt.dtype)
    t = t.contiguous()
    if tuple(t.shape) != tuple(tgt.shape):
        raise EqxvError(f"input shape {tuple(t.shape)} does not match the plan {tuple(tgt.shape)}")
    if t.is_cuda:
        if t.device != plan.device:
            raise EqxvError(f"input lives on {t.device}, the plan on {plan.device}")
        # produced on torch's current stream: the lane waits for it, the host does not
        _lib.call("eqxv_event_record", ctx.fence, torch.cuda.current_stream().cuda_stream)
        _lib.call("eqxv_stream_wait_event", plan.stream, ctx.fence)
    _lib.call("eqxv_memcpy_async", tgt.data_ptr(), t.data_ptr(), t.numel() * t.element_size(), plan.stream)
    if t.is_cuda:
        t.record_stream(torch.cuda.ExternalStream(plan.stream))


def run_batched(module, method: str, x, args=(), kwargs=None):
    """vmap(module.method)(x, *args, **kwargs): x is [N, ...per-sample shape] (fp32, the reference's input) or a
    transforms.ImagesU8 batch.

    Asynchronous like the reference's jit dispatch: the call enqueues copy -> graph -> output copy and returns CUDA
    tensors that are ordered on torch's current stream (use them with torch ops, `.cpu()`, or
    `eqxvision_b200.block_until_ready`). With host-resident inputs up to EQXV_LANES plan instances are used round-robin
    so that the copy of call i+1 overlaps the kernels of call i; a PINNED host batch must not be overwritten before the
    call that consumes it has finished (the usual non_blocking contract)."""
    from .transforms import ImagesU8

    ctx = _ctx()
    u8 = None
    if isinstance(x, ImagesU8):
        u8 = {"mean": x.mean, "std": x.std, "raw_hw": tuple(x.pixels.shape[1:3])}
        host = not x.pixels.is_cuda
    else:
        host = not (isinstance(x, torch.Tensor) and x.is_cuda)
    shape = tuple(x.shape)
    ps = get_plan_set(module, method, shape[0], shape[1:], args, kwargs, u8)
    ts = torch.cuda.current_stream().cuda_stream
    with ps.lock:
        plan = ps.pick(PIPELINE_LANES if host else 1)
        _enqueue_input(plan, x, ctx)
        if ctx.last_done is not None and ctx.last_done is not plan.done:
            _lib.call("eqxv_stream_wait_event", plan.stream, ctx.last_done)   # graphs run FIFO on the SMs
        plan.launch(plan.stream)
        _lib.call("eqxv_event_record", plan.done, plan.stream)
        ctx.last_done = plan.done
        # fresh result tensors (the reference returns new arrays): allocated by torch, written on the lane's stream
        # after everything already queued on torch's stream, visible to torch's stream once copied
        _lib.call("eqxv_event_record", ctx.fence, ts)
        _lib.call("eqxv_stream_wait_event", plan.stream, ctx.fence)
        leaves = []
        for o, shp in plan.outputs:
            r = torch.empty(o.shape, dtype=o.dtype, device=o.device)
            if o.is_contiguous():
                _lib.call("eqxv_memcpy_async", r.data_ptr(), o.data_ptr(), o.numel() * o.element_size(), plan.stream)
            else:   # logits in a buffer whose row pitch is padded to 8 columns
                ops.copy2d(r, o, stream=plan.stream)
            leaves.append(r.reshape(shp))
        _lib.call("eqxv_event_record", plan.out_ready, plan.stream)
        _lib.call("eqxv_stream_wait_event", ts, plan.out_ready)
    if BLOCKING:
        _lib.call("eqxv_stream_sync", plan.stream)
    return _unflatten_out(plan.out_struct, leaves)


def block_until_ready(out):
    """jax.block_until_ready for the tensors returned by a vmapped call"""
    torch.cuda.current_stream().synchronize()
    return out


def run_single(module, method: str, x, args=(), kwargs=None):
    """module.method(x) on ONE sample (the reference's native calling convention)"""
    if isinstance(x, torch.Tensor):
        xb = x.unsqueeze(0)
    else:
        xb = np.asarray(x, dtype=np.float32)[None]
    out = run_batched(module, method, xb, args, kwargs)

    def strip(o):
        if o is None:
            return None
        if isinstance(o, torch.Tensor):
            return o[0]
        return type(o)(strip(v) for v in o)

    return strip(out)
