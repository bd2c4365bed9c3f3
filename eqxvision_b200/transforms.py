"""The input edge of the hot path: what the reference's own fixture does to an image before the model sees it
(`/root/reference/tests/conftest.py:20-41`):

    transforms.Compose([transforms.Resize(img_size), transforms.ToTensor(),
                        transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])

In the reference this runs on the host (PIL + torch CPU) and the model receives an fp32 CHW array. Here the
uint8 HWC pixels are what travels to the GPU (a quarter of the bytes of the fp32 batch) and ToTensor + Normalize
are fused into the kernel that lays the image out for the first layer (`csrc/input_edge.cu`):

    batch = eqxvision_b200.transforms.images_u8(pixels_u8_nhwc)         # mean / std default to ImageNet's
    logits = eb.filter_jit(eb.vmap(net, axis_name="batch"))(batch, key=keys)

`ToTensor` + `Normalize` of a uint8 value has 256 outcomes per channel; `normalize_lut` evaluates them with the very
torch fp32 operations torchvision uses (`img.to(float32).div(255)`, `tensor.sub_(mean).div_(std)`), so the device
pipeline is bit-identical to the host pipeline up to the single bf16 rounding every activation gets.
`Resize` (optional): bilinear, half-pixel centres, no antialiasing - torchvision's tensor path
`F.interpolate(mode="bilinear", align_corners=False)`; PIL's antialiased filter is NOT reproduced.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_lut(mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """fp32 [C, 256]: lut[c][v] = Normalize(mean, std)(ToTensor(v)) with torchvision's operation order."""
    if len(mean) != len(std) or not 1 <= len(mean) <= 4:
        raise ValueError("mean / std must have one entry per channel (1..4 channels)")
    if any(s == 0 for s in std):
        raise ValueError("std evaluated to zero, leading to division by zero.")   # torchvision's message
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)           # ToTensor
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1)
    return v.view(1, 256).repeat(len(mean), 1).sub_(m).div_(s).contiguous()       # Normalize


class ImagesU8:
    """A batch of uint8 HWC images plus the transform the model input is derived from. Accepted by
    `vmap(net)(...)` in place of the fp32 NCHW batch; `shape` is the logical model-input shape (N, C, H, W)."""

    def __init__(self, pixels, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                 size: Optional[Tuple[int, int]] = None):
        if isinstance(pixels, np.ndarray):
            pixels = torch.from_numpy(np.ascontiguousarray(pixels))
        if not isinstance(pixels, torch.Tensor) or pixels.dtype != torch.uint8 or pixels.dim() != 4:
            raise TypeError("ImagesU8 expects a uint8 [N, H, W, C] array (PIL / numpy image layout)")
        if pixels.shape[3] != len(mean):
            raise ValueError(f"{pixels.shape[3]} channels but {len(mean)} mean / std entries")
        self.pixels = pixels.contiguous()
        self.mean = tuple(float(m) for m in mean)
        self.std = tuple(float(s) for s in std)
        if isinstance(size, int):
            size = (size, size)
        self.size = None if size is None else (int(size[0]), int(size[1]))
        normalize_lut(self.mean, self.std)   # validates

    @property
    def shape(self) -> Tuple[int, int, int, int]:
        n, h, w, c = self.pixels.shape
        if self.size is not None:
            h, w = self.size
        return (n, c, h, w)

    def pin_memory(self) -> "ImagesU8":
        return ImagesU8(self.pixels.pin_memory(), self.mean, self.std, self.size)

    def reference_pipeline(self) -> torch.Tensor:
        """the host pipeline of the reference fixture on the same pixels (fp32 NCHW), for tests: ToTensor + Normalize
        via torchvision's formulas; Resize through torch's bilinear interpolation in fp32 (see module docstring)"""
        x = self.pixels.cpu()
        if self.size is not None and tuple(x.shape[1:3]) != self.size:
            f = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2).float(), size=self.size, mode="bilinear",
                                                align_corners=False)
            x = f.round().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        t = x.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
        m = torch.as_tensor(self.mean, dtype=torch.float32).view(1, -1, 1, 1)
        s = torch.as_tensor(self.std, dtype=torch.float32).view(1, -1, 1, 1)
        return t.sub_(m).div_(s)


def images_u8(pixels, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
              size: Optional[Tuple[int, int]] = None) -> ImagesU8:
    return ImagesU8(pixels, mean, std, size)


def to_model_input(batch: ImagesU8, device=None) -> torch.Tensor:
    """Resize + ToTensor + Normalize on the device -> fp32 NCHW (what the reference pipeline hands to the model).
    The model call itself never materialises this tensor; it exists for callers that want the array."""
    from . import _engine, ops

    ctx = _engine._ctx()
    dev = ctx.device if device is None else torch.device(device)
    px = batch.pixels.to(dev)
    lut = normalize_lut(batch.mean, batch.std).to(dev)
    torch.cuda.current_stream().synchronize()
    if batch.size is not None and tuple(px.shape[1:3]) != batch.size:
        px = ops.u8_resize_bilinear(px, batch.size[0], batch.size[1], stream=ctx.stream)
    out = ops.u8_to_nchw_f32(px, lut, stream=ctx.stream)
    _engine._lib.call("eqxv_stream_sync", ctx.stream)
    return out
