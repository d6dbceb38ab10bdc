"""ctypes binding of ``libgnngp_b200.so`` (the C ABI declared in ``include/gnngp_b200.h``).

Prototypes are read from the header itself, so the binding cannot drift from the declared ABI and
``tests/test_abi.py`` can check that every declared symbol is exported.  There is no fallback: if the
library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "gnngp_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libgnngp_b200.so")

_SCALARS = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "size_t": ctypes.c_size_t,
    "float": ctypes.c_float, "double": ctypes.c_double, "gnngp_stream_t": ctypes.c_void_p,
}
_RETURNS = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "int64_t": ctypes.c_int64,
            "const char *": ctypes.c_char_p}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{symbol: (restype, [argtypes])} for every ``gnngp_*`` function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for match in re.finditer(r"(const char \*|size_t|int64_t|int)\s*(gnngp_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = match.group(1), match.group(2), match.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for arg in args.split(","):
                arg = arg.strip()
                if "*" in arg:
                    pname = arg.split("*")[-1].strip()
                    argtypes.append(ctypes.POINTER(ctypes.c_int) if pname.startswith("host_") else ctypes.c_void_p)
                else:
                    ctype = " ".join(arg.split()[:-1]).replace("const ", "").strip()
                    argtypes.append(_SCALARS[ctype])
        protos[name] = (_RETURNS[ret], argtypes)
    return protos


class Library(object):
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                "gnngp_b200: %s is missing. Build it with `python -m gnngp_b200.build` (needs nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for the GNNGP hot path." % path)
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)      # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        abi = self.cdll.gnngp_abi_version()
        if abi != 1:
            raise RuntimeError("gnngp_b200: ABI version %d, expected 1" % abi)

    def last_error(self) -> str:
        return (self.cdll.gnngp_last_error() or b"").decode("utf-8", "replace")

    def call(self, name: str, *args):
        """Call an int-returning entry point and raise on a non-zero status."""
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise RuntimeError("%s failed (status %d): %s" % (name, rc, self.last_error()))

    def query(self, name: str, *args):
        return getattr(self.cdll, name)(*args)


_LIBRARY = None


def library() -> Library:
    global _LIBRARY
    if _LIBRARY is None:
        _LIBRARY = Library()
        engine = os.environ.get("GNNGP_GEMM_ENGINE")
        if engine is not None:
            _LIBRARY.cdll.gnngp_set_gemm_engine(int(engine))
        flags = os.environ.get("GNNGP_SPMM_FLAGS")
        if flags is not None:
            _LIBRARY.cdll.gnngp_spmm_set_flags(int(flags))
    return _LIBRARY
