from .classification.resnet import (  # noqa: F401
    ResNet,
    resnet18,
    resnet34,
    resnet50,
    resnet101,
    resnet152,
    resnext50_32x4d,
    resnext101_32x8d,
    wide_resnet50_2,
    wide_resnet101_2,
)
from .classification.vit import (  # noqa: F401
    _VitAttention,
    _VitBlock,
    VisionTransformer,
    vit_base,
    vit_small,
    vit_tiny,
)
