#!/usr/bin/env python
"""Stage the UNMODIFIED reference next to the oracle so that it travels to the GPU box.

    python oracle/make_ref.py            (also called by __graft_entry__.build())

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is one pure-Python file on this path
(/root/reference/moegan/t2i_moe_gan.py).  It is copied byte for byte into ``oracle/_ref/`` - a directory that is
git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so the box that runs
``pytest -m gpu`` and ``bench.py`` receives it exactly like the built ``.so``.  Nothing is patched: the one missing
third-party import (``clip``, t2i_moe_gan.py:10 - OpenAI CLIP is not installable offline and nothing on the MoE path
calls it) is satisfied by an empty stub module that ``oracle/ref_loader.py`` registers before the import.

Consumers (all test / bench infrastructure): ``oracle/ref_loader.load_reference()`` ->
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg (kind "reference"),
  * ``bench.py``'s ``reference_gpu_eager`` rows (the same unmodified modules on the same B200),
  * ``tests/test_gpu_reference_dropin.py`` (``b200.install(ref)`` / ``SparseMoE.from_reference`` parity).
When ``/root/reference`` is absent (the GPU box) this script is a no-op and the staged copy is used as it is.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/moegan/t2i_moe_gan.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_DST = os.path.join(REF_DIR, "t2i_moe_gan.py")


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose: bool = True) -> str | None:
    """Copy the reference file into oracle/_ref/ (no-op without /root/reference).  Returns the staged path or None."""
    if not os.path.exists(REF_SRC):
        if verbose:
            print(f"make_ref: {REF_SRC} not present; {'using staged copy' if os.path.exists(REF_DST) else 'no reference available'}")
        return REF_DST if os.path.exists(REF_DST) else None
    os.makedirs(REF_DIR, exist_ok=True)
    if not os.path.exists(REF_DST) or sha256(REF_DST) != sha256(REF_SRC):
        shutil.copyfile(REF_SRC, REF_DST)
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": sha256(REF_DST), "bytes": os.path.getsize(REF_DST),
                   "note": "byte-for-byte staging copy of the unmodified reference file; git-ignored, test/bench infrastructure only"}, f, indent=1)
    if verbose:
        print(f"make_ref: staged {REF_DST} (sha256 {sha256(REF_DST)[:16]}...)")
    return REF_DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
