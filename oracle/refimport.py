"""Import the UNMODIFIED reference (veeresht/CommPy) read-only from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
oracle/validate_against_reference.py to pin the oracle and to generate tests/golden/.
The reference imports matplotlib at module top (convcode.py:12-17, modulation.py:25),
which is not installed, so empty stand-in modules are registered first.
"""
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("COMMPY_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "commpy"))


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


def import_reference():
    if not available():
        raise ImportError("reference checkout not found at %s" % REF_ROOT)
    for name in ("matplotlib", "matplotlib.colors", "matplotlib.patches", "matplotlib.path",
                 "matplotlib.pyplot", "matplotlib.collections"):
        if name not in sys.modules:
            sys.modules[name] = _Dummy(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import commpy  # noqa: F401
        import commpy.channelcoding  # noqa: F401
        import commpy.modulation  # noqa: F401
    return sys.modules["commpy"]
