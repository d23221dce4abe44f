"""ctypes loader of libmsl_frontend.so -- the C-ABI drop-in boundary (include/msl_frontend.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded the import of
any operator fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsl_frontend.so")
_lib = None


class MslError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("msl error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmsl_frontend.so is not built: run `python -m manhattanslam_b200.build` "
                              "(there is no CPU fallback for the CUDA front-end)")
        L = C.CDLL(LIB_PATH)
        L.msl_last_error.restype = C.c_char_p
        L.msl_version.restype = C.c_char_p
        L.msl_kernel_launch_count.restype = C.c_uint64
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MslError(rc, lib().msl_last_error().decode())


def ptr(a):
    """Raw pointer of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)
