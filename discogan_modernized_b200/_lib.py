"""ctypes binding of libdiscogan_b200.so.

The prototypes are read from ``include/discogan_b200.h`` so the header stays the single
source of truth for the C ABI.  There is no CPU fallback: if the library is missing the
first call raises.
"""
import ctypes
import os
import re
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent
HEADER = ROOT / "include" / "discogan_b200.h"
LIB_PATH = PKG_DIR / "libdiscogan_b200.so"
SOURCES = [PKG_DIR / "csrc" / n for n in ("gemm_tc.cu", "c3_tc.cu", "glue.cu", "direct.cu")]

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "size_t": ctypes.c_size_t, "dg_stream_t": ctypes.c_void_p,
}


class KernelError(RuntimeError):
    pass


def parse_header(text=None):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    text = HEADER.read_text() if text is None else text
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|size_t|long long)\s+(dg_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.groups()
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p,
                   "long long": ctypes.c_longlong}[ret]
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    continue
                ty = re.sub(r"\s+\w+$", "", a.replace("const ", "")).strip()
                argtypes.append(_CTYPES[ty])
        protos[name] = (restype, argtypes)
    return protos


def build(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a into the in-tree shared library (nvcc cross-compiles
    without a GPU)."""
    if LIB_PATH.exists() and not force:
        newest = max(p.stat().st_mtime for p in SOURCES + [PKG_DIR / "csrc" / "common.cuh", PKG_DIR / "csrc" / "tma_host.cuh"])
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-O3", "-std=c++17", "-o", str(LIB_PATH)] + [str(s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise KernelError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


_lib = None


def lib():
    """The loaded library with argtypes/restype set; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise KernelError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(L, name)  # AttributeError if the header declares something the .so lacks
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dg_last_error()
        raise KernelError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
