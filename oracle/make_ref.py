"""Recipe for oracle/_ref/: the reference's own model package, staged UNMODIFIED so that it travels to the GPU box
(TEST / BENCH INFRASTRUCTURE, never shipped: oracle/_ref/ is git-ignored and no product module imports it).

    python oracle/make_ref.py        # /root/reference/src/kp2dtiny -> oracle/_ref/src/kp2dtiny

The reference is pure Python (nothing to compile): "building" it is staging its `src.kp2dtiny` package -- models/,
modules/, utils/, 200 KB -- under the import path its own callers use (`from src.kp2dtiny.models.kp2dtiny import
tiny_factory`, eval_multitask.py:24).  `bench.py --impl reference` and the `cpu_baseline` leg import it from there
(`cpu_baseline.kind = "reference"`) and fall back to the oracle port only when the directory is absent.  Nothing is
written when /root/reference does not exist (the GPU box: it uses what travelled with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("NVS_REFERENCE", "/root/reference"), "src", "kp2dtiny")
DST = os.path.join(HERE, "_ref", "src", "kp2dtiny")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for d in (os.path.join(HERE, "_ref", "src"),):
        open(os.path.join(d, "__init__.py"), "a").close()  # `src` as a regular package next to the staged sub-package
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"[oracle] staged the reference model package: {SRC} -> {DST} ({n} files)")
    return True


def import_reference():
    """-> the reference's `src.kp2dtiny.models.kp2dtiny` module imported from oracle/_ref, or None when not staged."""
    root = os.path.join(HERE, "_ref")
    if not os.path.isdir(DST):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        if not getattr(sys.modules[name], "__file__", "") or root not in (sys.modules[name].__file__ or ""):
            del sys.modules[name]  # e.g. the product's compat shim registered under the same name
    return importlib.import_module("src.kp2dtiny.models.kp2dtiny")


if __name__ == "__main__":
    ok = make()
    sys.exit(0 if ok else 1)
