"""Classification model families; every submodule is importable as `models.classification.<family>`."""
import importlib

_FAMILIES = ("alexnet convnext densenet efficientnet googlenet mobilenetv2 mobilenetv3 regnet resnet shufflenetv2 "
             "squeezenet swin vgg vit").split()
for _name in _FAMILIES:
    globals()[_name] = importlib.import_module(f"{__name__}.{_name}")
del _name
