"""Stochastic depth (reference: layers/drop_path.py:8-61). Identity on the inference path."""
from .. import nn


class DropPath(nn.Module):
    p: float
    inference: bool
    mode: str

    def __init__(self, p: float = 0.0, inference: bool = False, mode="global"):
        self.p = p
        self.inference = inference
        self.mode = mode

    def __call__(self, x, *, key=None):
        if self.inference or self.p == 0.0:  # drop_path.py:44-45
            return x
        if key is None:
            raise RuntimeError(
                "DropPath requires a key when running in non-deterministic mode. Did you mean to enable inference?"
            )
        raise NotImplementedError(
            "training-mode DropPath is outside the inference hot path: use tree_inference(model, True)")
