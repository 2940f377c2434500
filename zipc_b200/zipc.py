"""Host-side mirror of the reference's `Zipc` module (src/zipc.mli) over the C ABI of libzipc_b200.so.

Only what sits on the hot path is mirrored: `File` (zipc.mli:100-199), `Member` (zipc.mli:203-262, the
parts that feed header bytes), `Ptime` DOS conversions, and the archive functions `of_binary_string`,
`encoding_size`, `to_binary_string` (zipc.mli:281-412) plus the batch forms a GPU needs:
`File.deflate_of_binary_strings`, `File.to_binary_strings` and `archive_of_binary_strings`.
Path pretty-printing, `Fpath.sanitize`, map folding helpers and the CLI are out of scope.

An archive is a dict {path: Member}; `add` replaces an existing path like the reference's String_map.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Iterable

import numpy as np

from . import _lib
from . import zipc_deflate as zd
from .zipc_deflate import Error, Ok

DOS_EPOCH = 315532800  # Ptime.dos_epoch, zipc.ml:94
MAX_SIZE = 0xFFFFFFFF   # File.max_size, zipc.ml:138
STORED, DEFLATE = 0, 8  # compression_to_int, zipc.ml:29-31
ZIP_REFERENCE, ZIP_ALLOW_ZIP64, ZIP_FORCE_ZIP64 = 0, 1, 2  # ZIPC_ZIP_* (include/zipc_b200.h)


class Ptime:
    dos_epoch = DOS_EPOCH

    @staticmethod
    def to_dos_date_time(ptime_s: int):
        d, t = C.c_int(), C.c_int()
        _lib.lib().zipc_b200_ptime_to_dos(ptime_s, C.byref(d), C.byref(t))
        return d.value, t.value

    @staticmethod
    def of_dos_date_time(dos_date: int, dos_time: int) -> int:
        return _lib.lib().zipc_b200_ptime_of_dos(dos_date, dos_time)


@dataclass
class FileT:
    """Zipc.File.t (zipc.ml:145-154)"""
    compression: int
    compressed_bytes: bytes | np.ndarray
    decompressed_size: int
    decompressed_crc_32: int
    start: int = 0
    compressed_size: int = 0
    version_made_by: int = (3 << 8) | 20
    version_needed_to_extract: int = 20
    gp_flags: int = 0x800


class File:
    max_size = MAX_SIZE

    @staticmethod
    def make(compressed_bytes, *, compression: int, decompressed_size: int, decompressed_crc_32: int, start: int = 0,
             compressed_size: int | None = None, version_made_by: int = (3 << 8) | 20,
             version_needed_to_extract: int = 20, gp_flags: int = 0x800):
        """zipc.ml:156-169"""
        if compressed_size is None:
            compressed_size = len(compressed_bytes) - start
        if compressed_size < 0 or decompressed_size < 0:
            raise ValueError("size is negative")  # Invalid_argument
        if start < 0 or start + compressed_size > len(compressed_bytes):
            # the reference fails later, in String.sub / the codec's bounds checks (Invalid_argument); the C side
            # trusts (start, compressed_size), so the range is checked where the record is made
            raise ValueError("index out of bounds")
        if compressed_size > MAX_SIZE or decompressed_size > MAX_SIZE:
            return Error("Maximum ZIP byte size 4294967295 exceeded by compressed (%d) or decompressed (%d) file size"
                         % (compressed_size, decompressed_size), 30)
        return Ok(FileT(compression, compressed_bytes, decompressed_size, decompressed_crc_32 & 0xFFFFFFFF, start,
                        compressed_size, version_made_by, version_needed_to_extract, gp_flags))

    @staticmethod
    def stored_of_binary_string(s, start: int = 0, len: int | None = None):
        """zipc.ml:171-177"""
        v = zd._slice(s, start, len)
        return File.make(v, compression=STORED, decompressed_size=v.size, decompressed_crc_32=zd.Crc_32.string(v))

    @staticmethod
    def deflate_of_binary_strings(items: Iterable, level: str = "default"):
        """Batch form of deflate_of_binary_string (zipc.ml:179-185): one GPU call for all payloads."""
        items = [zd._as_view(x) for x in items]
        out = []
        for v, (st, cs, crc) in zip(items, zd.default_context().deflate_batch(items, level, zd.CK_CRC32)):
            out.append(Error(zd.strerror(st), st) if st else
                       File.make(cs, compression=DEFLATE, decompressed_size=v.size, decompressed_crc_32=crc))
        return out

    @staticmethod
    def deflate_of_binary_string(s, level: str = "default", start: int = 0, len: int | None = None):
        return File.deflate_of_binary_strings([zd._slice(s, start, len)], level)[0]

    @staticmethod
    def is_encrypted(f: FileT) -> bool:
        return bool(f.gp_flags & 1)

    @staticmethod
    def can_extract(f: FileT) -> bool:
        return not File.is_encrypted(f) and f.compression in (STORED, DEFLATE)

    @staticmethod
    def to_binary_strings(files: list[FileT], check_crc: bool = True):
        """Batch form of to_binary_string[_no_crc_check] (zipc.ml:205-225)."""
        ms = [Member(path="", kind=f) for f in files]
        res = _extract(ms)
        out = []
        for f, (st, data, found) in zip(files, res):
            if st == _lib.ERR_CHECKSUM:
                out.append(Error(zd._crc_error(f.decompressed_crc_32, found), st, (f.decompressed_crc_32, found)) if check_crc
                           else Ok((data.tobytes(), found)))
            elif st in (_lib.ERR_CORRUPTED, _lib.ERR_SIZE_EXCEEDED):
                out.append(Error("deflate: " + zd.strerror(st), st))       # zipc.ml:215
            elif st == _lib.ERR_ZIP_FORMAT:
                out.append(Error("Compression %s not supported" % _compression_name(f.compression), st))
            elif st:
                out.append(Error(zd.strerror(st), st))
            else:
                out.append(Ok(data.tobytes()) if check_crc else Ok((data.tobytes(), found)))
        return out

    @staticmethod
    def to_binary_string(f: FileT):
        return File.to_binary_strings([f])[0]

    @staticmethod
    def to_binary_string_no_crc_check(f: FileT):
        return File.to_binary_strings([f], check_crc=False)[0]


def _compression_name(c: int) -> str:  # compression_to_string, zipc.ml:33-35
    return {12: "bz2", 8: "defl", 14: "lzma", 0: "none", 95: "xz", 93: "zst"}.get(c, "%04d" % c)


@dataclass
class Member:
    """Zipc.Member.t (zipc.ml:238-242); kind is None for a directory, else a FileT."""
    path: str | bytes
    kind: FileT | None
    mode: int = 0o644
    mtime: int = DOS_EPOCH

    max = 0xFFFF
    max_path_length = 0xFFFF

    @staticmethod
    def make(path, kind: FileT | None, mode: int | None = None, mtime: int = DOS_EPOCH):
        """zipc.ml:244-255"""
        p = path.encode() if isinstance(path, str) else bytes(path)
        p = p.replace(b"\\", b"/")
        if kind is None:
            p = b"./" if p == b"" else (p if p.endswith(b"/") else p + b"/")
        if len(p) > 0xFFFF:
            return Error("Maximum ZIP path length 65535 exceeded (%d)" % len(p), 29)
        if mode is None:
            mode = 0o755 if kind is None else 0o644
        return Ok(Member(p, kind, mode, max(mtime, DOS_EPOCH)))


def _to_c(members: list[Member]):
    n = len(members)
    arr = (_lib.Member * max(n, 1))()
    keep = []
    for i, m in enumerate(members):
        p = m.path.encode() if isinstance(m.path, str) else bytes(m.path)
        pb = C.create_string_buffer(p, len(p) or 1)
        keep.append(pb)
        f = m.kind
        if f is None:
            arr[i] = _lib.Member(C.cast(pb, C.c_void_p), len(p), 1, m.mode, m.mtime, 0, 0, 0, 0, None, 0, 0, 0, 0, 0)
        else:
            v = zd._as_view(f.compressed_bytes)
            if f.start < 0 or f.compressed_size < 0 or f.start + f.compressed_size > v.size:
                raise ValueError("index out of bounds")  # a hand-made FileT must not send the C side out of range
            keep.append(v)
            arr[i] = _lib.Member(C.cast(pb, C.c_void_p), len(p), 0, m.mode, m.mtime, f.version_made_by,
                                 f.version_needed_to_extract, f.gp_flags, f.compression, v.ctypes.data if v.size else None,
                                 f.start, f.compressed_size, f.decompressed_size, f.decompressed_crc_32, 0)
    return arr, keep


def _extract(members: list[Member]):
    ctx = zd.default_context()
    n = len(members)
    arr, _keep = _to_c(members)
    need = C.c_size_t()
    off, ln = (C.c_size_t * max(n, 1))(), (C.c_size_t * max(n, 1))()
    found, st = (C.c_uint32 * max(n, 1))(), (C.c_int * max(n, 1))()
    rc = ctx.L.zipc_b200_zip_extract_batch(ctx.h, arr, n, None, 0, C.byref(need), off, ln, found, st)
    if rc not in (_lib.OK, _lib.ERR_DST_TOO_SMALL):
        ctx._check(rc, "zip_extract_batch")
    arena = np.empty(max(need.value, 1), dtype=np.uint8)
    if rc == _lib.ERR_DST_TOO_SMALL:
        ctx._check(ctx.L.zipc_b200_fetch(ctx.h, arena.ctypes.data, arena.size), "fetch")
    return [(st[i], arena[off[i]:off[i] + ln[i]], found[i]) for i in range(n)]


# ---- archives ({path: Member}) -------------------------------------------------------------------------
def empty() -> dict:
    return {}


def add(m: Member, z: dict) -> dict:
    z = dict(z)
    z[bytes(m.path) if not isinstance(m.path, str) else m.path.encode()] = m
    return z


def find(path, z: dict):
    return z.get(path.encode() if isinstance(path, str) else bytes(path))


def member_count(z: dict) -> int:
    return len(z)


def string_has_magic(s) -> bool:
    """zipc.ml:427-430"""
    b = bytes(s[:4])
    return len(b) == 4 and b in (b"PK\x03\x04", b"PK\x05\x06")


def _zip_flags(zip64) -> int:
    """False: the reference's behaviour; True: ZIP64 structures are read / written where needed; "force": written always."""
    return ZIP_FORCE_ZIP64 | ZIP_ALLOW_ZIP64 if zip64 == "force" else ZIP_ALLOW_ZIP64 if zip64 else ZIP_REFERENCE


def of_binary_string(s, zip64: bool = False):
    """Zipc.of_binary_string (zipc.ml:432-438): members alias `s`.  zip64=True (beyond the reference, which answers
    "ZIP64 archives are not supported", zipc.ml:404) also reads archives with ZIP64 records."""
    v = zd._as_view(s)
    L = _lib.lib()
    p, n = C.POINTER(_lib.Member)(), C.c_size_t()
    st = L.zipc_b200_zip_parse_ex(v.ctypes.data if v.size else None, v.size, _zip_flags(zip64), C.byref(p), C.byref(n))
    if st:
        return Error(zd.strerror(st), st)
    z = {}
    for i in range(n.value):
        c = p[i]
        path = C.string_at(c.path, c.path_len)
        if c.is_dir:
            z[path] = Member(path, None, c.mode, c.mtime)
        else:
            z[path] = Member(path, FileT(c.compression, v, c.decompressed_size, c.crc32, c.start, c.compressed_size,
                                         c.version_made_by, c.version_needed, c.gp_flags), c.mode, c.mtime)
    L.zipc_b200_free(C.cast(p, C.c_void_p))
    return Ok(z)


def encoding_size(z: dict, zip64=False, first: str | bytes | None = None) -> int:
    """Zipc.encoding_size (zipc.ml:447-455); with zip64 the size of the ZIP64 records depends on the offsets, hence `first`."""
    arr, _keep = _to_c(list(z.values()))
    if not zip64:
        return _lib.lib().zipc_b200_zip_encoding_size(arr, len(z))
    f = first.encode() if isinstance(first, str) else first
    return _lib.lib().zipc_b200_zip_encoding_size_ex(arr, len(z), f, _zip_flags(zip64))


def to_binary_string(z: dict, first: str | bytes | None = None, zip64=False):
    """Zipc.to_binary_string (zipc.ml:585-588).  zip64=True (beyond the reference): more than 65,535 members, sizes and
    offsets of 4 GiB or more are written with ZIP64 records instead of being refused; an archive that needs none is
    byte-identical to the reference's.  zip64="force": ZIP64 records for every member."""
    ms = list(z.values())
    arr, _keep = _to_c(ms)
    L = _lib.lib()
    f = first.encode() if isinstance(first, str) else first
    flags = _zip_flags(zip64)
    size = L.zipc_b200_zip_encoding_size_ex(arr, len(ms), f, flags) or L.zipc_b200_zip_encoding_size(arr, len(ms))
    out = np.empty(max(size, 1), dtype=np.uint8)
    n = C.c_size_t()
    st = L.zipc_b200_zip_assemble_ex(arr, len(ms), f, flags, out.ctypes.data, size, C.byref(n))
    if st:
        return Error(zd.strerror(st), st)
    return Ok(out[:n.value].tobytes())


def archive_of_binary_strings(paths: list, payloads: list, level: str = "default", modes=None, mtimes=None,
                              first: str | bytes | None = None, zip64=False):
    """Batch form of File.deflate_of_binary_string + Member.make + add + to_binary_string: the payloads are
    compressed on the GPU, gathered to their archive offsets there, and the headers are laid around them."""
    ctx = zd.default_context()
    n = len(paths)
    pb = [p.encode() if isinstance(p, str) else bytes(p) for p in paths]
    vs = [zd._as_view(x) for x in payloads]
    pp = (C.c_void_p * max(n, 1))()
    keep = [C.create_string_buffer(p, len(p) or 1) for p in pb]
    pl = (C.c_uint32 * max(n, 1))()
    sp, sl = (C.c_void_p * max(n, 1))(), (C.c_size_t * max(n, 1))()
    for i in range(n):
        pp[i] = C.cast(keep[i], C.c_void_p).value
        pl[i] = len(pb[i])
        sp[i] = vs[i].ctypes.data if vs[i].size else None
        sl[i] = vs[i].size
    md = (C.c_int32 * n)(*modes) if modes is not None else None
    mt = (C.c_int64 * n)(*mtimes) if mtimes is not None else None
    f = first.encode() if isinstance(first, str) else first
    need = C.c_size_t()
    lv = zd.LEVELS[level]
    # generous first guess; the call reports the exact size when it does not fit
    cap = sum(v.size for v in vs) // 2 + (128 + (48 if zip64 else 0)) * n + 4096
    for _ in range(2):
        out = np.empty(max(cap, 1), dtype=np.uint8)
        st = ctx.L.zipc_b200_zip_deflate_archive_ex(ctx.h, lv, n, pp, pl, sp, sl, md, mt, f, _zip_flags(zip64), out.ctypes.data, cap, C.byref(need))
        if st == _lib.ERR_DST_TOO_SMALL:
            cap = need.value
            continue
        break
    if st:
        return Error(zd.strerror(st), st)
    return Ok(out[:need.value].tobytes())
