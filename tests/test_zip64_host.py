"""ZIP64 reading and writing (host only, no GPU): beyond the reference, which rejects ZIP64 archives
(zipc.ml:404) and refuses to write more than 65,535 members or 4 GiB fields (zipc.ml:229-235, 130-133, 550-551).
The reference's behaviour stays the default and is checked to be unchanged; the ZIP64 paths are checked against
CPython's zipfile in both directions (what it writes is parsed here, what is written here it reads)."""
import ctypes as C
import io
import struct
import zipfile
import zlib

import numpy as np
import pytest

from zipc_b200 import _lib, synth, zipc


def _raw_deflate(b: bytes) -> bytes:
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    return c.compress(b) + c.flush()


def _archive(n=5):
    z, plain = {}, {}
    for i in range(n):
        data = synth.text_v1(40 + i, 3000 + 997 * i).tobytes()
        path = b"dir/m%02d.txt" % i
        if i % 2:
            f = zipc.File.make(_raw_deflate(data), compression=zipc.DEFLATE, decompressed_size=len(data),
                               decompressed_crc_32=zlib.crc32(data)).get_ok()
        else:
            f = zipc.File.make(data, compression=zipc.STORED, decompressed_size=len(data), decompressed_crc_32=zlib.crc32(data)).get_ok()
        z[path] = zipc.Member.make(path, f, mtime=1_700_000_000 + i).get_ok()
        plain[path] = data
    z[b"dir/"] = zipc.Member.make(b"dir/", None).get_ok()
    return z, plain


def _same_members(a: dict, b: dict):
    assert sorted(a) == sorted(b)
    for k in a:
        ma, mb = a[k], b[k]
        assert (ma.mode, ma.mtime // 2, ma.kind is None) == (mb.mode, mb.mtime // 2, mb.kind is None)
        if ma.kind is not None:
            fa, fb = ma.kind, mb.kind
            assert (fa.compression, fa.decompressed_size, fa.decompressed_crc_32, fa.compressed_size) == \
                   (fb.compression, fb.decompressed_size, fb.decompressed_crc_32, fb.compressed_size)
            va, vb = np.frombuffer(fa.compressed_bytes, np.uint8), np.frombuffer(fb.compressed_bytes, np.uint8)
            assert bytes(va[fa.start:fa.start + fa.compressed_size]) == bytes(vb[fb.start:fb.start + fb.compressed_size])


def test_allowing_zip64_changes_nothing_when_it_is_not_needed(zip_docs):
    z, _ = _archive()
    ref = zipc.to_binary_string(z).get_ok()
    assert zipc.to_binary_string(z, zip64=True).get_ok() == ref
    assert zipc.encoding_size(z, zip64=True) == zipc.encoding_size(z) == len(ref)
    # the reference's own fixture: both parsers agree, and the re-encoding is the same bytes
    a, b = zipc.of_binary_string(zip_docs).get_ok(), zipc.of_binary_string(zip_docs, zip64=True).get_ok()
    _same_members(a, b)
    assert zipc.to_binary_string(a).get_ok() == zipc.to_binary_string(b, zip64=True).get_ok()


def test_forced_zip64_archive_is_read_by_cpython_and_by_the_parser():
    z, plain = _archive()
    forced = zipc.to_binary_string(z, zip64="force").get_ok()
    assert len(forced) == zipc.encoding_size(z, zip64="force")
    assert len(forced) > len(zipc.to_binary_string(z).get_ok())
    with zipfile.ZipFile(io.BytesIO(forced)) as zf:
        assert zf.testzip() is None
        assert sorted(i.filename.encode() for i in zf.infolist()) == sorted(z)
        for path, data in plain.items():
            assert zf.read(path.decode()) == data
            assert zf.getinfo(path.decode()).extract_version >= 45
    back = zipc.of_binary_string(forced, zip64=True).get_ok()
    _same_members(z, back)
    # ZIP64 end of central directory record + locator right in front of the classic record
    assert forced[-22 - 20:-22 - 16] == b"PK\x06\x07" and forced[-22 - 76:-22 - 72] == b"PK\x06\x06"


def test_more_than_65535_members():
    n = 66_000
    z = {}
    for i in range(n):
        p = b"e/%05d" % i
        z[p] = zipc.Member(p, zipc.FileT(zipc.STORED, b"", 0, 0))
    r = zipc.to_binary_string(z)
    assert not r.is_ok() and r.status == 28  # "Maximum ZIP member count 65535 exceeded (%d)", zipc.ml:231-232
    big = zipc.to_binary_string(z, zip64=True).get_ok()
    assert len(big) == zipc.encoding_size(z, zip64=True)
    with zipfile.ZipFile(io.BytesIO(big)) as zf:
        names = zf.namelist()
    assert len(names) == n and names[0] == "e/00000" and names[-1] == "e/%05d" % (n - 1)
    back = zipc.of_binary_string(big, zip64=True).get_ok()
    assert len(back) == n and list(back) == sorted(z)
    # made by CPython: the same number of members the other way round
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w") as zf:
        for i in range(n):
            zf.writestr("p/%05d" % i, b"")
    theirs = zipc.of_binary_string(bio.getvalue(), zip64=True).get_ok()
    assert len(theirs) == n


def test_sizes_of_4_gib_and_more_travel_in_the_extra_fields():
    five = 5 * (1 << 30)
    payload = _raw_deflate(b"\0" * 1000)  # (never inflated here: only the directory is under test)
    f = zipc.FileT(zipc.DEFLATE, payload, five, 0x12345678, 0, len(payload))
    z = {b"huge.bin": zipc.Member(b"huge.bin", f), b"small.txt": zipc.Member(b"small.txt", zipc.FileT(zipc.STORED, b"abc", 3, zlib.crc32(b"abc"), 0, 3))}
    r = zipc.to_binary_string(z)
    assert not r.is_ok()  # the reference cannot even make such a File (zipc.ml:130-133)
    s = zipc.to_binary_string(z, zip64=True).get_ok()
    assert len(s) == zipc.encoding_size(z, zip64=True)
    with zipfile.ZipFile(io.BytesIO(s)) as zf:
        info = zf.getinfo("huge.bin")
        assert info.file_size == five and info.compress_size == len(payload) and info.CRC == 0x12345678
        assert zf.read("small.txt") == b"abc"
        assert not any(x for x in zf.getinfo("small.txt").extra)  # only the member that needs it carries an extra field
    back = zipc.of_binary_string(s, zip64=True).get_ok()
    assert back[b"huge.bin"].kind.decompressed_size == five and back[b"huge.bin"].kind.compressed_size == len(payload)
    assert s[-22 - 20:-22 - 16] != b"PK\x06\x07"  # no ZIP64 end record: count and offsets fit


def test_offsets_of_4_gib_and_more_layout_only():
    """Three payloads of 3 GiB: the third local header and the directory start beyond 4 GiB.  Only the size
    computation runs (no buffer of that size is made): the C call is driven with sizes that are not backed by memory."""
    L = _lib.lib()
    arr = (_lib.Member * 3)()
    names = [C.create_string_buffer(b"a"), C.create_string_buffer(b"b"), C.create_string_buffer(b"c")]
    three = 3 * (1 << 30)
    for i in range(3):
        arr[i] = _lib.Member(C.cast(names[i], C.c_void_p), 1, 0, 0o644, zipc.DOS_EPOCH, 0x314, 20, 0x800, 0, C.cast(names[i], C.c_void_p),
                             0, three, three, 0, 0)
    assert L.zipc_b200_zip_encoding_size_ex(arr, 3, None, zipc.ZIP_REFERENCE) == 0   # "Maximum ZIP central directory offset ... exceeded"
    want = 3 * (30 + 1 + three) + 2 * (46 + 1) + (46 + 1 + 4 + 8) + 56 + 20 + 22  # one 8-byte offset in the third entry's extra field
    assert L.zipc_b200_zip_encoding_size_ex(arr, 3, None, zipc.ZIP_ALLOW_ZIP64) == want


def test_cpython_force_zip64_and_data_descriptors_parse():
    data = [synth.text_v1(7 + i, 20_000 + i).tobytes() for i in range(3)]
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w", zipfile.ZIP_DEFLATED) as zf:
        for i, d in enumerate(data):
            with zf.open("f%d.txt" % i, "w", force_zip64=True) as fh:  # local headers with ZIP64 extra fields
                fh.write(d)
    for zip64 in (False, True):  # the directory itself fits the classic format: the reference's parser reads it too
        z = zipc.of_binary_string(bio.getvalue(), zip64=zip64).get_ok()
        for i, d in enumerate(data):
            f = z[b"f%d.txt" % i].kind
            assert f.decompressed_size == len(d) and f.decompressed_crc_32 == zlib.crc32(d)
            raw = bytes(np.frombuffer(f.compressed_bytes, np.uint8)[f.start:f.start + f.compressed_size])
            assert zlib.decompress(raw, -15) == d

    class Unseekable(io.RawIOBase):  # makes zipfile write data descriptors (general purpose bit 3)
        def __init__(self): self.b = bytearray()
        def writable(self): return True
        def write(self, x): self.b += x; return len(x)
    u = Unseekable()
    with zipfile.ZipFile(u, "w", zipfile.ZIP_DEFLATED) as zf:
        for i, d in enumerate(data):
            zf.writestr("g%d.txt" % i, d)
    z = zipc.of_binary_string(bytes(u.b), zip64=True).get_ok()
    for i, d in enumerate(data):
        f = z[b"g%d.txt" % i].kind
        assert f.gp_flags & 8 and f.decompressed_size == len(d) and f.decompressed_crc_32 == zlib.crc32(d)  # sizes / CRC from the directory (zipc.ml:344-396)
        raw = bytes(np.frombuffer(f.compressed_bytes, np.uint8)[f.start:f.start + f.compressed_size])
        assert zlib.decompress(raw, -15) == d


def test_corrupt_zip64_records_are_refused():
    z, _ = _archive(2)
    del z[b"dir/"]  # (a directory entry's sizes are not looked at: zipc.ml:370-372)
    s = bytearray(zipc.to_binary_string(z, zip64="force").get_ok())
    loc = len(s) - 22 - 20
    bad = bytearray(s); struct.pack_into("<Q", bad, loc + 8, len(s) + 5)          # locator points outside
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 22
    bad = bytearray(s); struct.pack_into("<I", bad, loc + 16, 2)                   # two disks
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 21
    rec = loc - 56
    bad = bytearray(s); struct.pack_into("<QQ", bad, rec + 24, 1 << 40, 1 << 40)    # absurd member count
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 22
    bad = bytearray(s); struct.pack_into("<Q", bad, rec + 48, len(s))               # directory beyond the end
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 22
    cd = struct.unpack_from("<Q", s, rec + 48)[0]
    assert s[cd:cd + 4] == b"PK\x01\x02"
    plen = struct.unpack_from("<H", s, cd + 28)[0]
    bad = bytearray(s); struct.pack_into("<H", bad, cd + 46 + plen + 2, 8)          # extra field too short for three values
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 26
    bad = bytearray(s); struct.pack_into("<H", bad, cd + 46 + plen, 0x7075)         # no ZIP64 extra field at all
    assert zipc.of_binary_string(bytes(bad), zip64=True).status == 26
    # and the reference's answer to a ZIP64 marker in the classic record is kept without the flag
    bad = bytearray(s); bad[-18] = 0xFF; bad[-17] = 0xFF
    assert zipc.of_binary_string(bytes(bad)).status == 20
