"""Import shims that let the UNMODIFIED reference tree (python 3.6 / torch 1.7 / numpy 1.19 era) import under this
image's python 3.12 / numpy 2.  Test infrastructure only: used by tests/, bench.py's reference arm
and oracle/make_golden.py, never by the product package.

    src = ref_shims.ref_src()          # oracle/_ref/IntEL/src (copied there by oracle/make_ref.py) or /root/reference
    ref_shims.install(src)             # sys.path + `imp` + np.object / np.float / np.int
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref", "IntEL")


def ref_root() -> str | None:
    """The reference's IntEL directory: the copy that travels to the GPU box, else the read-only tree of the build container."""
    for p in (REF_COPY, "/root/reference/IntEL"):
        if os.path.isdir(os.path.join(p, "src", "models")):
            return p
    return None


def ref_src() -> str | None:
    r = ref_root()
    return os.path.join(r, "src") if r else None


def install(src: str) -> None:
    np.object = object                 # BaseModel.py:127,132
    np.float = float                   # utils.py:84
    np.int = int                       # utils.py:86
    if "imp" not in sys.modules:       # main.py:20 `from imp import reload`
        imp = types.ModuleType("imp")
        imp.reload = importlib.reload
        sys.modules["imp"] = imp
    if src not in sys.path:
        sys.path.insert(0, src)


def patch_predict_ragged(BaseRunner) -> None:
    """BaseRunner.predict calls np.array on per-session rows of different lengths and np.save()s them
    (BaseRunner.py:345-352): numpy >= 1.24 raises on the ragged list.  The patch builds object arrays instead; the
    returned values (what evaluate() consumes) are unchanged."""
    import numpy

    if getattr(BaseRunner, "_ragged_patched", False):
        return
    mod = sys.modules[BaseRunner.__module__]

    class _NP:
        def __getattr__(self, name):
            return getattr(numpy, name)

        @staticmethod
        def array(x, *a, **k):
            try:
                return numpy.array(x, *a, **k)
            except ValueError:
                out = numpy.empty(len(x), dtype=object)
                for i, v in enumerate(x):
                    out[i] = v
                return out

        @staticmethod
        def save(path, arr, *a, **k):
            return None                 # the .npy dumps next to the checkpoint are not part of the comparison

    mod.np = _NP()
    BaseRunner._ragged_patched = True
