"""Locate / load the product library. Fails loudly when it is missing: there is no Python or CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def find_lib_path():
    # GPB200_LIB: a variant build of the same library (gpboost_b200.build.build(extra_flags=..., out_name=...)), tuning runs only
    p = os.environ.get("GPB200_LIB") or os.path.join(_HERE, "lib_gpboost_b200.so")
    if not os.path.exists(p):
        raise RuntimeError("lib_gpboost_b200.so not found at %s — build it with `python -m gpboost_b200.build` "
                           "(or __graft_entry__.build()); there is no fallback implementation" % p)
    return p


def load_lib(path=None):
    """ctypes handle of the product library (cached), or of `path` when given (tests load the reference build
    of the same C API this way)."""
    global _LIB
    if path is not None:
        lib = ctypes.CDLL(path)
        lib.LGBM_GetLastError.restype = ctypes.c_char_p
        return lib
    if _LIB is None:
        _LIB = ctypes.CDLL(find_lib_path())
        _LIB.LGBM_GetLastError.restype = ctypes.c_char_p
        _LIB.gpbdev_last_error.restype = ctypes.c_char_p
        _LIB.gpbdev_vecchia_launch_count.restype = ctypes.c_int64
    return _LIB
