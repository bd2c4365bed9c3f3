"""ORACLE (test infrastructure) - import shim that runs the REFERENCE'S OWN model files on this box.

paganpasta/eqxvision is pure Python on top of jax / equinox, neither of which is (or can be) installed here. Its
arithmetic lives entirely in those third-party packages; its own code is the model wiring, the quirks and the
positional weight loader. `install()` puts minimal stand-ins for `jax`, `jax.numpy`, `jax.nn`, `jax.random`,
`jax.image`, `jax.lax`, `jax.tree_util`, `equinox`, `equinox.nn` and `equinox.experimental` into `sys.modules`
(numpy arrays with jax's 32-bit discipline, pytrees over dataclass-style modules, torch-CPU fp32 for conv / pooling /
resize) and imports the unmodified `eqxvision` package from /root/reference. The third-party semantics the stand-ins
encode are those of SURVEY.md 8(c)-S (Equinox 0.7-0.10 signatures and field order, JAX defaults).

What this pins: oracle/models.py (a hand restatement) against the reference's real control flow - constructors,
`load_torch_weights`, `__call__` - on identical checkpoints and inputs (tests/test_refshim.py), and the golden
vectors of tests/golden/golden_ref_v1.pt, which are outputs of the reference code executed this way. What it does not
pin: XLA's floating-point summation order (irrelevant at the 1e-4 tolerance the reference's own tests use).
Only tests and tests/golden/make_golden_ref.py import this package; it needs /root/reference and never travels to
the GPU box.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EQXV_REFERENCE_ROOT", "/root/reference")
_SHIMMED = ("jax", "jax.numpy", "jax.nn", "jax.random", "jax.image", "jax.lax", "jax.tree_util", "equinox",
            "equinox.nn", "equinox.experimental")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "eqxvision"))


@contextlib.contextmanager
def install():
    """with install() as eqxvision: ...  - the reference package, imported over the stand-ins"""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    from . import fake_equinox, fake_jax

    saved = {k: sys.modules.get(k) for k in _SHIMMED}
    saved_ref = {k: v for k, v in sys.modules.items() if k == "eqxvision" or k.startswith("eqxvision.")}
    for k in saved_ref:
        del sys.modules[k]
    mods = dict(fake_jax.modules())
    mods.update(fake_equinox.modules())
    sys.modules.update(mods)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield importlib.import_module("eqxvision")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "eqxvision" or k.startswith("eqxvision.")]:
            del sys.modules[k]
        sys.modules.update(saved_ref)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m
