"""ctypes binding of libdiscogan_b200.so.

The prototypes are read from ``include/discogan_b200.h`` so the header stays the single
source of truth for the C ABI.  There is no CPU fallback: if the library is missing the
first call raises.
"""
import ctypes
import hashlib
import os
import re
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent
HEADER = ROOT / "include" / "discogan_b200.h"
LIB_PATH = PKG_DIR / "libdiscogan_b200.so"
SOURCES = [PKG_DIR / "csrc" / n for n in ("gemm_tc.cu", "c3_tc.cu", "glue.cu", "direct.cu", "preprocess.cu")]

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


class ConvOpts(ctypes.Structure):
    """``dg_conv_opts`` of the header: per-call options of the tensor-core convolutions."""
    _fields_ = [("splitk_ws", ctypes.c_void_p), ("splitk_ws_bytes", ctypes.c_size_t), ("block_n", ctypes.c_int),
                ("pair", ctypes.c_int), ("wgrad_pair", ctypes.c_int), ("stat_accumulate", ctypes.c_int),
                ("affine_scale", ctypes.c_void_p), ("affine_shift", ctypes.c_void_p), ("affine_act", ctypes.c_int),
                ("affine_slope", ctypes.c_float)]


def source_hash():
    """sha256 over the CUDA sources, their shared headers and the C-ABI header (what the library is compiled from)."""
    h = hashlib.sha256()
    for p in SOURCES + [PKG_DIR / "csrc" / "common.cuh", PKG_DIR / "csrc" / "tma_host.cuh", HEADER]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:32]


def built_hash():
    """The source hash baked into the existing library (read from the file, without loading it), or None."""
    if not LIB_PATH.exists():
        return None
    m = re.search(rb"dgsrc:([0-9a-f]{32})", LIB_PATH.read_bytes())
    return m.group(1).decode() if m else None


def build(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU).
    The library records the hash of the sources it was built from (``dg_source_hash()``); it is rebuilt whenever
    that hash differs from the sources on disk, or when ``force`` / ``DG_FORCE_BUILD=1`` asks for it."""
    want = source_hash()
    force = force or os.environ.get("DG_FORCE_BUILD", "0") == "1"
    if not force and built_hash() == want:
        if verbose:
            print(f"{LIB_PATH.name}: up to date (sources {want})")
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-O3", "-std=c++17", f"-I{HEADER.parent}", f'-DDG_SOURCE_HASH="dgsrc:{want}"', "-o", str(tmp)] + [str(s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise KernelError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
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
        got = L.dg_source_hash().decode()
        if got != "dgsrc:" + source_hash() and os.environ.get("DG_ALLOW_STALE_LIB", "0") != "1":
            raise KernelError(f"{LIB_PATH.name} was built from other sources ({got}) than the ones on disk "
                              f"(dgsrc:{source_hash()}): rebuild with `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dg_last_error()
        raise KernelError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
