"""Import unmodified reference files that ``import modal`` at module level (scripts/*.py, demo/server/server.py).

``modal`` is not in this image.  A stub module whose attributes swallow any call and whose decorators return the
decorated object is put into ``sys.modules`` first; nothing of the reference is edited.  Build container only."""
import importlib.util
import os
import sys
import types

REF = "/root/reference"


class _Stub:
    """Stands in for every attribute of ``modal``: calling it with one plain callable (a decorator use) returns that
    callable, anything else returns another stub; usable as a context manager."""

    def __getattr__(self, k):
        return _Stub()

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Stub):
            return a[0]
        return _Stub()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def load_reference_file(relpath: str, module_name: str):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "modal" not in sys.modules:
        modal = types.ModuleType("modal")

        def _attr(name):                                  # PEP 562: any public attribute is a stub
            if name.startswith("__"):
                raise AttributeError(name)
            return _Stub()
        modal.__getattr__ = _attr
        sys.modules["modal"] = modal
    spec = importlib.util.spec_from_file_location(module_name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
