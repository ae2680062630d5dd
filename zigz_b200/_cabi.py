"""ctypes loader for libzigz_b200.so (device C ABI zb_* + host twin zh_*).

The prototypes are PARSED from include/zigz_b200.h and include/zigz_host.h, so the Python view can never drift
from the declared boundary.  There is no fallback of any kind: a missing library raises, a missing GPU makes
zb_ctx_create return ZB_ERR_NO_DEVICE which `check()` turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libzigz_b200.so")
HEADERS = [os.path.join(ROOT, "include", "zigz_b200.h"), os.path.join(ROOT, "include", "zigz_host.h")]

_CTYPES = {
    "int32_t": C.c_int32, "int64_t": C.c_int64, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "size_t": C.c_size_t, "int": C.c_int,
    "uint8_t": C.c_uint8, "zb_mle": C.c_uint64, "zb_tree": C.c_uint64, "void": None, "char": C.c_char, "float": C.c_float,
    "double": C.c_double, "zb_rank_fn": C.c_void_p,  # callback pointer (host twins only)
}
_OPAQUE = {"zb_ctx", "zh_transcript"}


def _ctype(decl: str):
    """C parameter / return declaration -> ctypes type."""
    d = decl.replace("const", " ").strip()
    d = re.sub(r"\[[^\]]*\]", "*", d)  # arrays decay to pointers
    stars = d.count("*")
    d = d.replace("*", " ")
    toks = d.split()
    base = toks[0] if toks else "void"
    if base == "struct":
        base = toks[1]
    if base in _OPAQUE:
        return C.c_void_p  # any pointer depth to an opaque struct travels as void*
    t = _CTYPES[base]
    if stars == 0:
        return t
    if base in ("void",):
        return C.c_void_p
    if base == "char":
        return C.c_char_p if stars == 1 else C.c_void_p
    for _ in range(stars):
        t = C.POINTER(t)
    return t


_PROTO = re.compile(r"^\s*(?!typedef|return)((?:const\s+)?[A-Za-z_][\w]*(?:\s*\*+)?)\s*(z[bh]_\w+)\s*\(([^;{]*?)\)\s*;", re.M | re.S)


def declared_prototypes():
    """[(name, restype_decl, [param_decl, ...])] for every zb_*/zh_* function the headers declare."""
    out = []
    for h in HEADERS:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
        src = re.sub(r"//[^\n]*", " ", src)
        for m in _PROTO.finditer(src):
            ret, name, params = m.group(1), m.group(2), m.group(3)
            plist = [p.strip() for p in params.split(",")] if params.strip() and params.strip() != "void" else []
            out.append((name, ret.strip(), plist))
    return out


def build(force: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (make is incremental)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        subprocess.run(args + ["clean"], check=True, capture_output=True)
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libzigz_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, ret, params in declared_prototypes():
            fn = getattr(L, name)  # AttributeError if the header declares something the library lacks
            fn.restype = _ctype(ret)
            fn.argtypes = [_ctype(p) for p in params]
        _lib = L
    return _lib


class ZigzError(Exception):
    """A Zig error name of the reference (error.NoVariables, ...) or a device error."""

    def __init__(self, code: int, detail: str = ""):
        self.code = code
        self.name = lib().zb_status_name(code).decode()
        super().__init__(f"error.{self.name}" + (f" ({detail})" if detail else ""))
