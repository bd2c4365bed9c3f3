from .deeplabv3 import ASPP, ASPPConv, ASPPPooling, DeepLabHead, DeepLabV3, deeplabv3  # noqa: F401
from .fcn import FCN, FCNHead, fcn  # noqa: F401
from .lraspp import LRASPP, LRASPPHead, lraspp_mobilenet_v3_large  # noqa: F401
