"""Segmentation wrapper (reference: models/segmentation/_utils.py:8-60): backbone taps -> classifier on
the last tap -> bilinear resize to the input resolution; the same for the auxiliary head on the
first tap; returns `(aux, out)` (aux first, _utils.py:58)."""
from typing import Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom


class _SimpleSegmentationModel(nn.Module):
    backbone: nn.Module
    classifier: nn.Module
    aux_classifier: nn.Module

    def __init__(self, backbone: nn.Module, classifier: nn.Module, aux_classifier: Optional[nn.Module] = None) -> None:
        self.backbone = backbone
        self.classifier = classifier
        self.aux_classifier = aux_classifier

    def __call__(self, x, *, key=None):
        k1, k2, k3 = jrandom.split(key, 3)
        h, w = x.shape[-2:]
        _, taps = self.backbone(x, key=k1)
        out = F.resize_bilinear(self.classifier(taps[-1], key=k2), h, w)
        if self.aux_classifier is None:
            return None, out
        aux = F.resize_bilinear(self.aux_classifier(taps[0], key=k3), h, w)
        return aux, out
