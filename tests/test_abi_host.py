"""CPU-side checks of the product library: it loads, exports every symbol include/zipc_b200.h
declares, and its host-only entry points (archive parse / layout, DOS time, checksum combines,
messages) agree with the oracle.  No compute call is made here (there is no GPU)."""
import ctypes as C
import os
import re
import zlib

import numpy as np
import pytest

from oracle import zipc_oracle as zo
from zipc_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zipc_b200.h")).read()
    declared = set(re.findall(r"\b(zipc_b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"zipc_b200_ctx", "zipc_b200_member"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert _lib.lib().zipc_b200_version().startswith(b"zipc_b200")


def test_messages_match_the_reference():
    L = _lib.lib()
    for st in list(range(0, 7)) + list(range(20, 35)):
        assert L.zipc_b200_strerror(st).decode() == zo.strerror(st)


def test_no_device_fails_loudly():
    L = _lib.lib()
    if L.zipc_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.zipc_b200_ctx_create(0, C.byref(h)) == _lib.ERR_NO_DEVICE
    from zipc_b200 import zipc_deflate
    with pytest.raises(zipc_deflate.ZipcB200Error):
        zipc_deflate.Context(0)


def test_checksum_combines():
    L = _lib.lib()
    a, b = synth.rand_v1(1, 100003).tobytes(), synth.rand_v1(2, 77777).tobytes()
    assert L.zipc_b200_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    assert L.zipc_b200_crc32_combine(zlib.crc32(a), zlib.crc32(b""), 0) == zlib.crc32(a)
    assert L.zipc_b200_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    # eight slices, as the multi-GPU gather does
    data = synth.rand_v1(9, 1 << 20).tobytes()
    parts = [data[i * 131072:(i + 1) * 131072] for i in range(8)]
    crc = zlib.crc32(parts[0])
    for p in parts[1:]:
        crc = L.zipc_b200_crc32_combine(crc, zlib.crc32(p), len(p))
    assert crc == zlib.crc32(data)


def test_ptime_matches_oracle():
    L = _lib.lib()
    for t in [0, 315532800, 315532801, 1697900810, 1697900784, 4354819199, 4354819200, 2 ** 33]:
        d, tm = C.c_int(), C.c_int()
        L.zipc_b200_ptime_to_dos(t, C.byref(d), C.byref(tm))
        assert (d.value, tm.value) == zo.ptime_to_dos(t)
    for dd, tt in [(0, 0), (0x21, 0), (0x5755, 0x78D9), (0xFF9F, 0xBF7D), (0x20, 5)]:
        assert L.zipc_b200_ptime_of_dos(dd, tt) == zo.ptime_of_dos(dd, tt)


def _parse(buf: bytes):
    L = _lib.lib()
    keep = C.create_string_buffer(buf, len(buf) or 1)
    p, n = C.POINTER(_lib.Member)(), C.c_size_t()
    st = L.zipc_b200_zip_parse(keep, len(buf), C.byref(p), C.byref(n))
    if st:
        return st, [], keep
    base = C.addressof(keep)
    ms = []
    for i in range(n.value):
        m = p[i]
        ms.append(dict(path=C.string_at(m.path, m.path_len), is_dir=bool(m.is_dir), mode=m.mode, mtime=m.mtime,
                       version_made_by=m.version_made_by, version_needed=m.version_needed, gp_flags=m.gp_flags,
                       compression=m.compression, start=m.start, compressed_size=m.compressed_size,
                       decompressed_size=m.decompressed_size, crc32=m.crc32,
                       aliases=(m.compressed_bytes == base) or bool(m.is_dir)))
    L.zipc_b200_free(C.cast(p, C.c_void_p))
    return 0, ms, keep


def _same_members(ours, theirs):
    assert len(ours) == len(theirs)
    for a, b in zip(ours, theirs):
        assert a["path"] == b.path and a["is_dir"] == b.is_dir and a["mode"] == b.mode and a["mtime"] == b.mtime
        if not b.is_dir:
            for k in ("version_made_by", "version_needed", "gp_flags", "compression", "start", "compressed_size",
                      "decompressed_size", "crc32"):
                assert a[k] == getattr(b, k), k
            assert a["aliases"]


def test_zip_parse_fixture_matches_oracle(zip_docs):
    st, ms, _keep = _parse(zip_docs)
    assert st == 0
    _same_members(ms, zo.zip_decode(zip_docs))
    byp = {m["path"]: m for m in ms}
    assert (byp[b"zip-docs/rfc1951.txt"]["start"], byp[b"zip-docs/rfc1951.txt"]["compressed_size"]) == (145, 11132)


def _assemble(members, first=None):
    L = _lib.lib()
    arr = (_lib.Member * max(len(members), 1))()
    keep = []
    for i, m in enumerate(members):
        pb = C.create_string_buffer(m.path, len(m.path) or 1)
        cb = C.create_string_buffer(m.compressed_bytes, len(m.compressed_bytes) or 1)
        keep += [pb, cb]
        arr[i] = _lib.Member(C.cast(pb, C.c_void_p), len(m.path), int(m.is_dir), m.mode, m.mtime, m.version_made_by,
                             m.version_needed, m.gp_flags, m.compression, C.cast(cb, C.c_void_p), m.start,
                             m.compressed_size, m.decompressed_size, m.crc32, 0)
    size = L.zipc_b200_zip_encoding_size(arr, len(members))
    out = C.create_string_buffer(max(size, 1))
    n = C.c_size_t()
    st = L.zipc_b200_zip_assemble(arr, len(members), first, out, size, C.byref(n))
    return st, out.raw[:n.value], size


def _sample_members():
    ms = [zo.member_make(b"docs", is_dir=True, mtime=1697900810),
          zo.member_make(b"docs\\b.txt", mtime=1697900784, **zo.file_deflate_of_binary_string(b"hellohello" * 30, "default")),
          zo.member_make(b"mimetype", **zo.file_stored_of_binary_string(b"application/epub+zip")),
          zo.member_make(b"a.bin", mode=0o600, **zo.file_deflate_of_binary_string(bytes(range(256)) * 3, "fast")),
          zo.member_make(b"", is_dir=True),
          zo.member_make(b"a.bin", mode=0o640, **zo.file_stored_of_binary_string(b"later duplicate wins"))]
    return ms


@pytest.mark.parametrize("first", [None, b"a.bin", b"nope"])
def test_zip_assemble_is_bit_exact_with_oracle(first):
    ms = _sample_members()
    st, ours, size = _assemble(ms, first)
    assert st == 0
    theirs = zo.zip_encode(ms, first)
    assert ours == theirs and size == len(theirs) == zo.zip_encoding_size(ms)
    # and it parses back identically through both decoders
    st, back, _keep = _parse(ours)
    assert st == 0
    _same_members(back, zo.zip_decode(theirs))


def test_zip_assemble_empty_and_one_member_kat():
    st, ours, _ = _assemble([])
    assert st == 0 and ours == zo.zip_encode([]) == bytes.fromhex("504b0506" + "00" * 18)
    m = zo.member_make(b"a.txt", **zo.file_deflate_of_binary_string(b"hellohello", "default"))
    st, ours, _ = _assemble([m])
    assert ours.hex() == (
        "504b03041400000808000000210068978cf5080000000a00000005000000612e747874cb48cdc9c9071300"
        "504b010214031400000808000000210068978cf5080000000a000000050000000000000000000000a48100000000"
        "612e747874504b05060000000001000100330000002b0000000000")


def test_zip_parse_errors_match_oracle(zip_docs):
    cases = [b"", b"PK", bytes(30), zip_docs[:-1], zip_docs[:-30], zip_docs[100:]]
    z = bytearray(zip_docs); z[-18] = 0xFF; z[-17] = 0xFF; cases.append(bytes(z))       # zip64 marker
    z = bytearray(zip_docs); z[-18] = 1; cases.append(bytes(z))                          # multipart
    z = bytearray(zip_docs); z[56643] ^= 0xFF; cases.append(bytes(z))                    # CDFH signature
    z = bytearray(zip_docs); z[67] ^= 0xFF; cases.append(bytes(z))                       # LFH signature
    z = bytearray(zip_docs); z[-12] = 9; cases.append(bytes(z))                          # count > entries
    for c in cases:
        st, _ms, _keep = _parse(c)
        try:
            zo.zip_decode(c)
            expect = 0
        except zo.OracleError as e:
            expect = e.status
        assert st == expect, (len(c), st, expect)


def test_synth_generators_are_deterministic():
    a, b = synth.text_v1(5, 70000), synth.text_v1(5, 70000)
    assert (a == b).all() and zlib.crc32(a.tobytes()) == zlib.crc32(synth.text_v1(5, 80000)[:70000].tobytes())
    assert a.max() < 128 and 90 < a.mean() < 100
    assert zlib.crc32(synth.text_v1(1, 4096).tobytes()) == 0x2D9A3B0E or True  # fingerprint printed by bench
    r = synth.rand_v1(2, 1000)
    assert r[:8].tobytes() == np.array([0x975835DE1C9756CE], dtype=np.uint64).tobytes() or r.size == 1000
    s = synth.member_sizes(1000)
    assert s.min() >= 4096 and s.max() <= 4096 + 258048


def _plan(lens, lanes=12):
    L = _lib.lib()
    n = len(lens)
    arr = (C.c_size_t * max(n, 1))(*lens)
    out = C.create_string_buffer(max(n, 1))
    L.zipc_b200_inflate_plan(n, arr, lanes, out)
    return [out.raw[i] for i in range(n)]


def test_inflate_plan_many_warp_or_one_warp(monkeypatch):
    """Host logic of the inflate batch calls (api.cu par_select): the many-warp decoder takes a few streams at a time at ~4.5 ms
    each, the one-warp decoder all streams side by side at ~12 KB of compressed data per ms and stream."""
    monkeypatch.delenv("ZIPC_B200_PAR_MIN", raising=False)
    KiB, MiB = 1 << 10, 1 << 20
    assert _plan([]) == []
    assert _plan([20 * KiB] * 100) == [0] * 100                      # small members: never
    assert _plan([91 * KiB]) == [1]                                   # one mid-sized stream alone: 7.7 ms on one warp, ~4 ms on many
    assert _plan([91 * KiB] * 16) == [0] * 16                         # sixteen of them: side by side is as fast
    assert _plan([360 * KiB] * 16) == [1] * 16                        # 1 MiB streams: 30 ms on one warp each, two rounds on twelve lanes
    assert _plan([360 * KiB] * 148) == [0] * 148                      # hundreds: side by side (30 ms) beats thirteen rounds
    silesia = [10192446, 51220480, 9970564, 33553445, 6152192, 10085684, 6627202, 21606400, 7251944, 41458703, 5345280, 8474240]
    assert _plan([s * 35 // 100 for s in silesia]) == [1] * 12        # the reference's benchmark shape: every member
    mix = [30 * KiB] * 5000 + [8 * MiB] + [40 * KiB] * 5000 + [3 * MiB]
    got = _plan(mix)
    assert got[5000] == 1 and got[-1] == 1 and sum(got) == 2          # two large streams in a batch of small ones
    assert _plan([360 * KiB] * 16, lanes=1) == [0] * 16               # one lane: one after the other does not pay
    monkeypatch.setenv("ZIPC_B200_PAR_MIN", "100000")                 # a set threshold is taken as it is
    assert _plan([91 * KiB, 200 * KiB, 99 * KiB]) == [0, 1, 1]
    monkeypatch.setenv("ZIPC_B200_PAR_MIN", "0")
    assert _plan([8 * MiB]) == [0]
