"""Import the UNMODIFIED reference from /root/reference (build container only).

The reference cannot be imported as-is here: `src/models.py:10-11` pulls in
`faiss` (src/database/faiss_store.py:12), `objectbox`
(src/database/objectbox_store.py:12, entities.py) and `matplotlib`
(src/utils.py:3), none of which are installed, and its default constructor
downloads GPT-2 (`src/models.py:211`, `src/utils.py:100`).  We pre-insert inert
stub modules into `sys.modules` and use the constructor's own injection points
(`gpt=`, `tokenizer=`, src/models.py:183-191) -- nothing under /root/reference is
modified or copied.

This module is used ONLY by `tests/golden/make_golden.py` and by the
"reference present" tests; it does not exist on the GPU box (no /root/reference
there), where everything runs against the committed fixtures instead.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GIC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models.py"))


class _Anything:
    """Callable / subscriptable / attribute-able placeholder for stubbed symbols."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        # used both as a plain call and as a decorator factory (objectbox.Entity())
        if len(a) == 1 and isinstance(a[0], type) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub_module(name: str, attrs: tuple[str, ...] = ()) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__["__stub__"] = True
    for a in attrs:
        setattr(mod, a, _Anything())

    def _getattr(attr):  # any other symbol the reference may touch at import time
        if attr.startswith("__"):  # keep `inspect` & co. honest (__file__, __path__, ...)
            raise AttributeError(attr)
        return _Anything()

    mod.__getattr__ = _getattr  # type: ignore[attr-defined]
    return mod


def install_shims() -> None:
    if "faiss" not in sys.modules:
        sys.modules["faiss"] = _stub_module("faiss")
    if "objectbox" not in sys.modules:
        sys.modules["objectbox"] = _stub_module(
            "objectbox",
            ("Box", "Model", "Store", "Entity", "Float32Vector", "Float64", "HnswIndex",
             "Id", "Int64", "String", "VectorDistanceType"),
        )
    if "matplotlib" not in sys.modules:
        mpl = _stub_module("matplotlib")
        plt = _stub_module("matplotlib.pyplot")
        mpl.pyplot = plt  # type: ignore[attr-defined]
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def import_reference():
    """Returns (src.models, src.database.faiss_store) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    models = importlib.import_module("src.models")
    fstore = importlib.import_module("src.database.faiss_store")
    return models, fstore


class StubTokenizer:
    """Only `eos_token_id` is read on the greedy path (src/models.py:348)."""

    eos_token_id = 50256
