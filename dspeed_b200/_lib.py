"""Loader (and builder) of the C-ABI CUDA library ``libdspeed_b200.so``.

The library is built IN-TREE with nvcc for sm_100a (``build()``; called by
``__graft_entry__.build``) and loaded with ctypes.  There is no CPU fallback: if the
library is missing or cannot be loaded, importing any processor raises."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.path.join(_HERE, "libdspeed_b200.so")
SOURCES = ["processors.cu", "conv.cu", "fused.cu", "conv_tc.cu", "sipm.cu", "glue.cu"]
HEADERS = ["common.cuh", "row_ops.cuh", "conv_ops.cuh", "conv_seg.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
]

_lock = threading.Lock()
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, "dspeed_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libdspeed_b200.so`` (in-tree)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-o", LIB_PATH, *sources()]
    if verbose:
        print(" ".join(cmd))
    env = dict(os.environ)
    # nvcc must use the system host compiler (the image's $CC wrapper lacks some specs)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Raises (loudly) when it is not built: there is no fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  dspeed_b200 has no CPU fallback."
                    )
                _lib = C.CDLL(LIB_PATH)
                _lib.dspb_fatal_message.restype = C.c_char_p
                _lib.dspb_max_row_len.restype = C.c_int64
    return _lib


def fatal_message(code: int) -> str:
    return lib().dspb_fatal_message(int(code)).decode()
