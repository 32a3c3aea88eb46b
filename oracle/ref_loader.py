"""Import the UNMODIFIED reference module (moegan/t2i_moe_gan.py) as test / bench infrastructure.

TEST / BENCH INFRASTRUCTURE ONLY - nothing under ``moe-gan_cpsc541_b200/`` imports this file.

Search order: ``oracle/_ref/t2i_moe_gan.py`` (the byte-for-byte staging copy made by ``oracle/make_ref.py`` at build
time; this is what exists on the GPU box), then ``/root/reference/moegan/t2i_moe_gan.py`` (the build container).
The module does ``import clip`` at file scope (t2i_moe_gan.py:10); OpenAI CLIP is not installable offline and is never
called on the MoE path, so an empty stub module is registered first.  Nothing else is patched.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref", "t2i_moe_gan.py"), "/root/reference/moegan/t2i_moe_gan.py")
_MOD = None


def reference_path():
    for p in CANDIDATES:
        if os.path.exists(p):
            return p
    return None


def available() -> bool:
    return reference_path() is not None


def load_reference():
    """The reference module object (cached).  Raises FileNotFoundError when neither copy exists."""
    global _MOD
    if _MOD is not None:
        return _MOD
    path = reference_path()
    if path is None:
        raise FileNotFoundError("unmodified reference not found: run `python oracle/make_ref.py` in the build container "
                                "(stages /root/reference/moegan/t2i_moe_gan.py into oracle/_ref/)")
    sys.modules.setdefault("clip", types.ModuleType("clip"))
    spec = importlib.util.spec_from_file_location("t2i_moe_gan_reference", path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    mod.__reference_path__ = path
    _MOD = mod
    return mod


def make_loss(ref, device):
    """AuroraGANLoss without CLIP: __init__ would call clip.load (t2i_moe_gan.py:915); generator_loss,
    discriminator_loss and moe_balance_loss never touch clip_loss_fn."""
    obj = ref.AuroraGANLoss.__new__(ref.AuroraGANLoss)
    obj.device = device
    return obj
