"""eqxvision_b200 — B200-native implementation of eqxvision's vmapped inference forward pass.

Drop-in surface (reference: paganpasta/eqxvision 0.2.8):
    eqxvision_b200.models.*      constructors (`resnet50(torch_weights=...)`, `vit_base(...)`, ...)
    eqxvision_b200.layers.*      ConvNormActivation, SqueezeExcitation, PatchEmbed, MlpProjection, ...
    eqxvision_b200.utils         load_torch_weights, CLASSIFICATION_URLS, SEGMENTATION_URLS
    eqxvision_b200.vmap / filter_jit / tree_inference   stand-ins for jax.vmap / eqx.filter_jit /
                                 eqx.tree_inference so that README-style call sites run unchanged.
    eqxvision_b200.transforms    the uint8 input edge (Resize / ToTensor / Normalize of the reference's fixture, on the GPU)
All arithmetic runs in libeqxv_b200.so (hand-written sm_100a CUDA behind a C ABI); there is no CPU
or PyTorch fallback.
"""
__version__ = "0.2.8+b200.1"

from . import experimental, layers, models, nn, random, transforms, utils  # noqa: F401
from ._engine import block_until_ready  # noqa: F401
from .compat import filter_jit, tree_inference, vmap  # noqa: F401
