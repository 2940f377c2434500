"""ctypes front end of the parity oracle (oracle/libzipc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under zipc_b200/ imports this module.

Function names mirror the reference modules they restate (Zipc_deflate.Crc_32.string ->
crc32, Zipc_deflate.inflate -> inflate, Zipc.to_binary_string -> zip_encode ...); each returns
either the value or raises OracleError carrying the reference's status + message.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzipc_oracle.so")

OK, ERR_CORRUPTED, ERR_SIZE_EXCEEDED = 0, 1, 2
ERR_ZLIB_METHOD, ERR_ZLIB_WINDOW, ERR_ZLIB_DICT, ERR_CHECKSUM = 3, 4, 5, 6
CRC_NOP, CRC_ADLER32, CRC_CRC32 = 0, 1, 2
LEVELS = {"none": 0, "fast": 1, "default": 2, "best": 3}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile) if it is missing or stale."""
    src = os.path.join(_HERE, "zipc_oracle.c")
    hdr = os.path.join(_HERE, "zipc_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


class OracleError(Exception):
    def __init__(self, status: int, message: str, **extra):
        super().__init__(message)
        self.status, self.message, self.extra = status, message, extra


class _Member(C.Structure):
    _fields_ = [("path", C.c_void_p), ("path_len", C.c_uint32), ("is_dir", C.c_int),
                ("mode", C.c_int), ("mtime", C.c_int64),
                ("version_made_by", C.c_int), ("version_needed", C.c_int), ("gp_flags", C.c_int),
                ("compression", C.c_int), ("compressed_bytes", C.c_void_p),
                ("start", C.c_uint64), ("compressed_size", C.c_uint64),
                ("decompressed_size", C.c_uint64), ("crc32", C.c_uint32)]


class _Stats(C.Structure):
    _fields_ = [("blocks_stored", C.c_uint32), ("blocks_fixed", C.c_uint32),
                ("blocks_dynamic", C.c_uint32), ("literals", C.c_uint64),
                ("matches", C.c_uint64), ("match_bytes", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, vpp, szp, u32p = C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)
        L.zo_strerror.restype = C.c_char_p
        L.zo_crc32.restype = C.c_uint32
        L.zo_crc32.argtypes = [u8p, C.c_size_t]
        L.zo_crc32_update.restype = C.c_uint32
        L.zo_crc32_update.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.zo_adler32.restype = C.c_uint32
        L.zo_adler32.argtypes = [u8p, C.c_size_t]
        L.zo_adler32_update.restype = C.c_uint32
        L.zo_adler32_update.argtypes = [C.c_uint32, u8p, C.c_size_t]
        L.zo_inflate.argtypes = [u8p, C.c_size_t, C.c_int64, C.c_int, vpp, szp, u32p]
        L.zo_zlib_decompress.argtypes = [u8p, C.c_size_t, C.c_int64, vpp, szp, u32p, u32p, u32p,
                                         C.POINTER(C.c_int)]
        L.zo_deflate.argtypes = [C.c_int, u8p, C.c_size_t, C.c_int, vpp, szp, u32p, C.POINTER(_Stats)]
        L.zo_zlib_compress.argtypes = [C.c_int, u8p, C.c_size_t, vpp, szp, u32p]
        L.zo_free.argtypes = [C.c_void_p]
        L.zo_ptime_of_dos.restype = C.c_int64
        L.zo_ptime_of_dos.argtypes = [C.c_int, C.c_int]
        L.zo_ptime_to_dos.argtypes = [C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.zo_ptime_to_date_time.argtypes = [C.c_int64] + [C.POINTER(C.c_int)] * 6
        L.zo_zip_encode.argtypes = [C.POINTER(_Member), C.c_size_t, C.c_char_p, vpp, szp]
        L.zo_zip_encoding_size.restype = C.c_uint64
        L.zo_zip_encoding_size.argtypes = [C.POINTER(_Member), C.c_size_t]
        L.zo_zip_decode.argtypes = [u8p, C.c_size_t, C.POINTER(C.POINTER(_Member)), szp]
        L.zo_file_to_binary_string.argtypes = [C.POINTER(_Member), vpp, szp, u32p]
        L.zo_member_make_path.restype = C.c_void_p
        L.zo_member_make_path.argtypes = [C.c_char_p, C.c_size_t, C.c_int, szp]
        _lib = L
    return _lib


def strerror(status: int) -> str:
    return lib().zo_strerror(status).decode()


def set_adler_signed_rem(on: bool) -> None:
    lib().zo_set_adler_signed_rem(int(on))


def set_keep_codelen_freqs(on: bool) -> None:
    lib().zo_set_keep_codelen_freqs(int(on))


def _take(ptr: C.c_void_p, n: int) -> bytes:
    data = C.string_at(ptr.value, n) if n else b""
    lib().zo_free(ptr)
    return data


# ---- Zipc_deflate.Crc_32 / Adler_32 -----------------------------------------------------------
def crc32(s: bytes) -> int:
    return lib().zo_crc32(s, len(s))


def crc32_update(c: int, s: bytes) -> int:
    return lib().zo_crc32_update(c, s, len(s))


def adler32(s: bytes) -> int:
    return lib().zo_adler32(s, len(s))


def adler32_update(a: int, s: bytes) -> int:
    return lib().zo_adler32_update(a, s, len(s))


# ---- Zipc_deflate.inflate* ---------------------------------------------------------------------
def inflate_and_crc(s: bytes, decompressed_size: int | None = None, crc_op: int = CRC_NOP):
    out, n, crc = C.c_void_p(), C.c_size_t(), C.c_uint32()
    st = lib().zo_inflate(s, len(s), -1 if decompressed_size is None else decompressed_size,
                          crc_op, C.byref(out), C.byref(n), C.byref(crc))
    if st:
        raise OracleError(st, strerror(st))
    return _take(out, n.value), crc.value


def inflate(s: bytes, decompressed_size: int | None = None) -> bytes:
    return inflate_and_crc(s, decompressed_size, CRC_NOP)[0]


def inflate_and_crc_32(s: bytes, decompressed_size: int | None = None):
    return inflate_and_crc(s, decompressed_size, CRC_CRC32)


def inflate_and_adler_32(s: bytes, decompressed_size: int | None = None):
    return inflate_and_crc(s, decompressed_size, CRC_ADLER32)


def zlib_decompress(s: bytes, decompressed_size: int | None = None):
    out, n, ad = C.c_void_p(), C.c_size_t(), C.c_uint32()
    ex, fo, me = C.c_uint32(), C.c_uint32(), C.c_int()
    st = lib().zo_zlib_decompress(s, len(s), -1 if decompressed_size is None else decompressed_size,
                                  C.byref(out), C.byref(n), C.byref(ad), C.byref(ex), C.byref(fo),
                                  C.byref(me))
    if st == ERR_CHECKSUM:
        raise OracleError(st, "Checksum mismatch, expected %x found %x)" % (ex.value, fo.value),
                          expect=ex.value, found=fo.value)
    if st == ERR_ZLIB_METHOD:
        raise OracleError(st, "Unknown compression method (%d)" % me.value)
    if st:
        raise OracleError(st, strerror(st))
    return _take(out, n.value), ad.value


# ---- Zipc_deflate.deflate* ---------------------------------------------------------------------
def crc_and_deflate(s: bytes, level: str = "best", crc_op: int = CRC_NOP, stats: dict | None = None):
    """NB the reference's omitted ?level is `Best (zipc_deflate.ml:817), hence the default."""
    out, n, crc, stt = C.c_void_p(), C.c_size_t(), C.c_uint32(), _Stats()
    st = lib().zo_deflate(LEVELS[level], s, len(s), crc_op, C.byref(out), C.byref(n), C.byref(crc),
                          C.byref(stt))
    if st:
        raise OracleError(st, strerror(st))
    if stats is not None:
        stats.update({k: getattr(stt, k) for k, _ in _Stats._fields_})
    return crc.value, _take(out, n.value)


def deflate(s: bytes, level: str = "best", stats: dict | None = None) -> bytes:
    return crc_and_deflate(s, level, CRC_NOP, stats)[1]


def crc_32_and_deflate(s: bytes, level: str = "best"):
    return crc_and_deflate(s, level, CRC_CRC32)


def adler_32_and_deflate(s: bytes, level: str = "best"):
    return crc_and_deflate(s, level, CRC_ADLER32)


def zlib_compress(s: bytes, level: str = "best"):
    out, n, ad = C.c_void_p(), C.c_size_t(), C.c_uint32()
    st = lib().zo_zlib_compress(LEVELS[level], s, len(s), C.byref(out), C.byref(n), C.byref(ad))
    if st:
        raise OracleError(st, strerror(st))
    return ad.value, _take(out, n.value)


# ---- Zipc.Ptime ------------------------------------------------------------------------------
DOS_EPOCH = 315532800


def ptime_of_dos(dos_date: int, dos_time: int) -> int:
    return lib().zo_ptime_of_dos(dos_date, dos_time)


def ptime_to_dos(t: int):
    d, tm = C.c_int(), C.c_int()
    lib().zo_ptime_to_dos(t, C.byref(d), C.byref(tm))
    return d.value, tm.value


def ptime_to_date_time(t: int):
    v = [C.c_int() for _ in range(6)]
    lib().zo_ptime_to_date_time(t, *[C.byref(x) for x in v])
    return (v[0].value, v[1].value, v[2].value), (v[3].value, v[4].value, v[5].value)


# ---- Zipc.Member / Zipc.File / archive ---------------------------------------------------------
@dataclass
class Member:
    """Member.t + File.t (zipc.ml:145-154, 238-242)."""
    path: bytes
    is_dir: bool = False
    mode: int = 0o644
    mtime: int = DOS_EPOCH
    version_made_by: int = 0x314
    version_needed: int = 20
    gp_flags: int = 0x800
    compression: int = 8
    compressed_bytes: bytes = b""
    start: int = 0
    compressed_size: int = 0
    decompressed_size: int = 0
    crc32: int = 0
    _keep: list = field(default_factory=list, repr=False, compare=False)


def member_make(path: bytes, *, is_dir: bool = False, mode: int | None = None,
                mtime: int = DOS_EPOCH, **file_fields) -> Member:
    """Member.make (zipc.ml:244-255) + File.make defaults (zipc.ml:138-169)."""
    n = C.c_size_t()
    p = lib().zo_member_make_path(path, len(path), int(is_dir), C.byref(n))
    norm = C.string_at(p, n.value)
    lib().zo_free(p)
    if len(norm) > 0xFFFF:
        raise OracleError(29, "Maximum ZIP path length 65535 exceeded (%d)" % len(norm))
    if mode is None:
        mode = 0o755 if is_dir else 0o644
    m = Member(path=norm, is_dir=is_dir, mode=mode, mtime=max(mtime, DOS_EPOCH), **file_fields)
    if not is_dir and "compressed_size" not in file_fields:
        m.compressed_size = len(m.compressed_bytes) - m.start
    return m


def _to_c(ms: list[Member]):
    arr = (_Member * max(len(ms), 1))()
    keep = []
    for i, m in enumerate(ms):
        pb = C.create_string_buffer(m.path, len(m.path) or 1)
        cb = C.create_string_buffer(m.compressed_bytes, len(m.compressed_bytes) or 1)
        keep += [pb, cb]
        arr[i] = _Member(C.cast(pb, C.c_void_p), len(m.path), int(m.is_dir), m.mode, m.mtime,
                         m.version_made_by, m.version_needed, m.gp_flags, m.compression,
                         C.cast(cb, C.c_void_p), m.start, m.compressed_size, m.decompressed_size,
                         m.crc32)
    return arr, keep


def zip_encoding_size(ms: list[Member]) -> int:
    arr, _keep = _to_c(ms)
    return lib().zo_zip_encoding_size(arr, len(ms))


def zip_encode(ms: list[Member], first: bytes | None = None) -> bytes:
    """Zipc.to_binary_string (zipc.ml:585-588)."""
    arr, _keep = _to_c(ms)
    out, n = C.c_void_p(), C.c_size_t()
    st = lib().zo_zip_encode(arr, len(ms), first, C.byref(out), C.byref(n))
    if st:
        raise OracleError(st, strerror(st))
    return _take(out, n.value)


def zip_decode(s: bytes) -> list[Member]:
    """Zipc.of_binary_string (zipc.ml:432-438): members sorted by path."""
    buf = C.create_string_buffer(s, len(s) or 1)
    base = C.addressof(buf)
    p, n = C.POINTER(_Member)(), C.c_size_t()
    st = lib().zo_zip_decode(C.cast(buf, C.c_char_p), len(s), C.byref(p), C.byref(n))
    if st:
        raise OracleError(st, strerror(st))
    res = []
    for i in range(n.value):
        c = p[i]
        path = C.string_at(c.path, c.path_len)
        if c.is_dir:
            res.append(Member(path=path, is_dir=True, mode=c.mode, mtime=c.mtime))
        else:
            assert c.compressed_bytes == base
            res.append(Member(path=path, is_dir=False, mode=c.mode, mtime=c.mtime,
                              version_made_by=c.version_made_by, version_needed=c.version_needed,
                              gp_flags=c.gp_flags, compression=c.compression, compressed_bytes=s,
                              start=c.start, compressed_size=c.compressed_size,
                              decompressed_size=c.decompressed_size, crc32=c.crc32))
    lib().zo_free(C.cast(p, C.c_void_p))
    return res


def file_to_binary_string(m: Member) -> bytes:
    """Zipc.File.to_binary_string (zipc.ml:219-225)."""
    arr, _keep = _to_c([m])
    out, n, found = C.c_void_p(), C.c_size_t(), C.c_uint32()
    st = lib().zo_file_to_binary_string(arr, C.byref(out), C.byref(n), C.byref(found))
    if st == ERR_CHECKSUM:
        raise OracleError(st, "Checksum mismatch, expected %x found %x)" % (m.crc32, found.value),
                          expect=m.crc32, found=found.value)
    if st in (ERR_CORRUPTED, ERR_SIZE_EXCEEDED):
        raise OracleError(st, "deflate: " + strerror(st))
    if st:
        raise OracleError(st, strerror(st))
    return _take(out, n.value)


def file_deflate_of_binary_string(s: bytes, level: str = "best") -> dict:
    """Zipc.File.deflate_of_binary_string (zipc.ml:179-185): fields for member_make."""
    crc, cs = crc_32_and_deflate(s, level)
    return dict(compression=8, compressed_bytes=cs, start=0, compressed_size=len(cs),
                decompressed_size=len(s), crc32=crc)


def file_stored_of_binary_string(s: bytes) -> dict:
    """Zipc.File.stored_of_binary_string (zipc.ml:171-177)."""
    return dict(compression=0, compressed_bytes=s, start=0, compressed_size=len(s),
                decompressed_size=len(s), crc32=crc32(s))
