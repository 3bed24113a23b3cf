"""Loader of ``libw2t.so`` (the CUDA library behind ``include/w2t.h``).

There is no CPU fallback: if the library has not been built, or no CUDA device
is visible when a compute entry point is called, the product raises.
"""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# W2T_LIB: an experimental build of the same library (debug aid for A/B timing, csrc/Makefile)
LIB_PATH = os.environ.get("W2T_LIB") or os.path.join(_HERE, "libw2t.so")
_lib = None


class W2TError(RuntimeError):
    pass


def lib():
    """The bound library; raises W2TError when it is missing (never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise W2TError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C waymo_2d_tracking_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        _lib = _abi.bind(ctypes.CDLL(LIB_PATH))
    return _lib


def check(status, what):
    if status != _abi.W2T_OK:
        msg = lib().w2t_last_error().decode() or _abi.STATUS_NAMES.get(status, "status %d" % status)
        lib().w2t_clear_error()              # a later failure without a message of its own must not show this one
        raise W2TError("%s failed: %s" % (what, msg))


def check_device_status(code, what):
    if code != _abi.W2T_OK:
        raise W2TError("%s: kernel reported %s" % (what, _abi.STATUS_NAMES.get(int(code), "status %d" % code)))


def build(verbose=False):
    """Compile libw2t.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    import subprocess
    res = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise W2TError("nvcc build of libw2t.so failed")
    return LIB_PATH
