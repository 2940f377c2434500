"""Loader of the host-side encoder model (tools/deflate_model.cc): test scaffolding that drives the
same __host__ __device__ building blocks as the CUDA encoder, serially."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tools", "libdeflate_model.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "tools", "deflate_model.cc")
        core = os.path.join(ROOT, "zipc_b200", "csrc", "deflate_core.h")
        if not os.path.exists(_SO) or max(os.path.getmtime(src), os.path.getmtime(core)) > os.path.getmtime(_SO):
            subprocess.run([os.path.join(ROOT, "tools", "build_model.sh")], check=True)
        L = C.CDLL(_SO)
        L.zipc_model_deflate.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def deflate(data: bytes, level: str) -> bytes:
    o, n = C.c_void_p(), C.c_uint64()
    lib().zipc_model_deflate({"fast": 1, "default": 2, "best": 3}[level], data, len(data), C.byref(o), C.byref(n))
    out = C.string_at(o.value, n.value)
    lib().zipc_model_free(o)
    return out
