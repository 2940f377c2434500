"""Pins the CPU oracle (oracle/zipc_oracle.c) before anything is compared against it.

Sources of truth, in order of authority:
  1. the reference's own test vectors           /root/reference/test/test.ml:14-129 + fixture
  2. known answers of the survey-time independent emulation of zipc_deflate.ml (SURVEY.md 8c,
     BASELINE.md 3b) -- byte-exact agreement of two independent restatements
  3. system zlib / zipfile as a second oracle for inflate, CRC-32, RFC Adler-32, readability
"""
import io
import random
import zipfile
import zlib

import pytest

from oracle import zipc_oracle as zo

FOX = b"The quick brown fox jumps over the lazy dog"
LEVELS = ("none", "fast", "default", "best")


# ---- test/test.ml:16-26 -------------------------------------------------------------------------
def test_crc_32_reference_vectors():
    assert zo.crc32(b"") == 0
    assert zo.crc32(FOX) == 0x414FA339


def test_adler_32_reference_vectors():
    assert zo.adler32(b"") == 1
    assert zo.adler32(FOX) == 0x5BDC0FDA


# ---- test/test.ml:28-43 -------------------------------------------------------------------------
TRIP = [
    (b"", "fixed"),
    (b"a", "fixed"),
    (b"hellohello", "fixed"),
    (b"abcdefghijklmnopqrstuvwxyzzyxwvutsrqponmlkjihgfedcba", "dynamic"),
    (bytes((i + 1) % 255 for i in range(256)), "stored"),
]


@pytest.mark.parametrize("level", LEVELS)
@pytest.mark.parametrize("s,kind", TRIP)
def test_deflate_trip(level, s, kind):
    st = {}
    cs = zo.deflate(s, level, st)
    assert zo.inflate(cs) == s
    assert zlib.decompress(cs, -15) == s
    if level == "none":
        kind = "stored"
    # the block kinds the reference's comments state (test.ml:38-42)
    assert st["blocks_" + kind] == 1 and sum(st["blocks_" + k] for k in ("stored", "fixed", "dynamic")) == 1


# ---- test/test.ml:45-55 -------------------------------------------------------------------------
def test_decompression_size_limits():
    src = b"Keep it to the limits."
    csrc = zo.deflate(src)
    assert zo.inflate(csrc) == src
    assert zo.inflate(csrc, len(src)) == src
    assert zo.inflate(csrc, len(src) + 1) == src
    with pytest.raises(zo.OracleError) as e:
        zo.inflate(csrc, len(src) - 1)
    assert e.value.message == "Expected decompression size exceeded"


# ---- test/test.ml:57-118 ------------------------------------------------------------------------
def _assert_zip(archive, original):
    ms = {m.path: m for m in zo.zip_decode(archive)}
    d = ms[b"zip-docs/"]
    assert d.is_dir and zo.ptime_to_date_time(d.mtime) == ((2023, 10, 21), (15, 6, 50)) and d.mode == 0o755
    r = ms[b"zip-docs/rfc1951.txt"]
    assert zo.ptime_to_date_time(r.mtime) == ((2023, 10, 21), (15, 6, 24)) and r.mode == 0o644
    a = ms[b"zip-docs/APPNOTE.TXT"]
    assert zo.ptime_to_date_time(a.mtime) == ((2023, 10, 21), (15, 6, 50)) and a.mode == 0o644
    if original:
        assert r.compression == 8 and r.decompressed_size == 36944
        assert a.compression == 8 and a.decompressed_size == 174585
    assert r.crc32 == 0xFB4F3400 and a.crc32 == 0x39B029C4
    return ms, zo.file_to_binary_string(r), zo.file_to_binary_string(a)


def test_crunched_trip(zip_docs):
    ms, r, a = _assert_zip(zip_docs, True)
    # second oracle: zlib / zipfile agree on the payloads
    zf = zipfile.ZipFile(io.BytesIO(zip_docs))
    assert zf.read("zip-docs/rfc1951.txt") == r and zf.read("zip-docs/APPNOTE.TXT") == a
    # redeflate_recode (test.ml:58-73): omitted ?level, i.e. `Best
    out = []
    for m in ms.values():
        if m.is_dir:
            out.append(m)
        else:
            s = zo.file_to_binary_string(m)
            out.append(zo.member_make(m.path, mode=m.mode, mtime=m.mtime,
                                      **zo.file_deflate_of_binary_string(s)))
    recoded = zo.zip_encode(out)
    assert len(recoded) == zo.zip_encoding_size(out)
    _ms2, r2, a2 = _assert_zip(recoded, True)
    assert (r2, a2) == (r, a)
    zf2 = zipfile.ZipFile(io.BytesIO(recoded))
    assert zf2.testzip() is None and zf2.read("zip-docs/APPNOTE.TXT") == a


def test_fixture_decode_details(zip_docs):
    # data offsets from the LFH's own name+extra lengths (zipc.ml:332-334), SURVEY.md 4
    ms = {m.path: m for m in zo.zip_decode(zip_docs)}
    assert (ms[b"zip-docs/rfc1951.txt"].start, ms[b"zip-docs/rfc1951.txt"].compressed_size) == (145, 11132)
    assert (ms[b"zip-docs/APPNOTE.TXT"].start, ms[b"zip-docs/APPNOTE.TXT"].compressed_size) == (11355, 45288)
    assert ms[b"zip-docs/rfc1951.txt"].version_made_by == 0x031E


# ---- survey-time emulation KATs (SURVEY.md 8c / BASELINE.md 3b) -----------------------------------
ADLER_QUIRK = [  # (data, as written in the reference, RFC 1950)
    (b"\xff" * 5552, 0xF0BD9B8C, 0xF18F9B8C),
    (b"\xff" * 5553, 0x8D579C8B, 0x8E299C8B),
    (b"\xff" * 100000, 0x04D7302C, 0x149A302C),
    (bytes(range(256)) * 1000, 0x9142292F, 0xA73B292F),
    (b"\x7f" * 100000, 0xB039D4B0, 0xB741D4B0),
    (b"a" * 10 ** 6, 0x15D870F9, 0x15D870F9),
]


@pytest.mark.parametrize("data,as_written,rfc", ADLER_QUIRK)
def test_adler_signed_rem_quirk(data, as_written, rfc):
    assert zo.adler32(data) == as_written
    assert zlib.adler32(data) == rfc
    zo.set_adler_signed_rem(False)
    try:
        assert zo.adler32(data) == rfc
    finally:
        zo.set_adler_signed_rem(True)


def test_deflate_small_kats():
    assert zo.deflate(b"").hex() == "0300"
    assert zo.deflate(b"a").hex() == "4b0400"
    assert zo.deflate(b"hellohello").hex() == "cb48cdc9c9071300"
    for lvl in ("fast", "default", "best"):
        assert zo.deflate(b"Keep it to the limits.", lvl).hex() == \
            "f34e4d2d50c82c5128c95728c94855c8c9cccd2c29d60300"
    assert zo.deflate(b"", "none").hex() == "010000ffff"


REDEFLATE = {  # (size, CRC-32 of the produced stream) of the reference algorithm
    ("r", "fast"): (11557, 0x5AFF2DA8), ("r", "default"): (11114, 0x2C8EFD2F), ("r", "best"): (11100, 0x7DDC7CCF),
    ("a", "fast"): (47961, 0x0816A253), ("a", "default"): (45224, 0x96DE8297), ("a", "best"): (45097, 0xFC510ACE),
    ("ra", "fast"): (59126, 0x1E7AC9A9), ("ra", "default"): (55851, 0x32341D22), ("ra", "best"): (55701, 0xE8CF941F),
}


@pytest.mark.parametrize("key", sorted(REDEFLATE))
def test_fixture_redeflate_fingerprints(zip_docs, key):
    ms = {m.path: m for m in zo.zip_decode(zip_docs)}
    r = zo.file_to_binary_string(ms[b"zip-docs/rfc1951.txt"])
    a = zo.file_to_binary_string(ms[b"zip-docs/APPNOTE.TXT"])
    data = {"r": r, "a": a, "ra": r + a}[key[0]]
    out = zo.deflate(data, key[1])
    assert (len(out), zlib.crc32(out)) == REDEFLATE[key]
    assert zlib.decompress(out, -15) == data and zo.inflate(out) == data
    if key[0] == "ra":
        assert zlib.crc32(data) == 0x09A56728 and zlib.adler32(data) == 0xC669A4BA
        assert zo.crc32(data) == 0x09A56728 and zo.adler32(data) == 0xC669A4BA


def test_one_member_archive_kat():
    m = zo.member_make(b"a.txt", **zo.file_deflate_of_binary_string(b"hellohello", "default"))
    assert zo.zip_encode([m]).hex() == (
        "504b03041400000808000000210068978cf5080000000a00000005000000612e747874cb48cdc9c9071300"
        "504b010214031400000808000000210068978cf5080000000a000000050000000000000000000000a48100000000"
        "612e747874504b05060000000001000100330000002b0000000000")


# ---- second oracle: zlib -------------------------------------------------------------------------
def _corpus():
    rnd = random.Random(7)
    words = [bytes(rnd.choice(b"etaoinshrdlu") for _ in range(rnd.randint(2, 9))) for _ in range(300)]
    text = b" ".join(rnd.choice(words) for _ in range(40000))
    return {
        "text": text,
        "random": rnd.randbytes(100000),
        "zeros": bytes(70000),
        "mixed": text[:50000] + rnd.randbytes(30000) + text[:50000],
        "short": b"abc",
    }


@pytest.mark.parametrize("name", ["text", "random", "zeros", "mixed", "short"])
def test_inflate_of_zlib_streams(name):
    data = _corpus()[name]
    for lvl in (0, 1, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY):
            c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strategy)
            cs = c.compress(data) + c.flush()
            out, crc = zo.inflate_and_crc_32(cs, len(data))
            assert out == data and crc == zlib.crc32(data)
            out, ad = zo.inflate_and_adler_32(cs)
            assert out == data


@pytest.mark.parametrize("name", ["text", "random", "zeros", "mixed", "short"])
@pytest.mark.parametrize("level", LEVELS)
def test_deflate_inflates_with_zlib(name, level):
    data = _corpus()[name]
    crc, cs = zo.crc_32_and_deflate(data, level)
    assert crc == zlib.crc32(data)
    assert zlib.decompress(cs, -15) == data
    ad, zs = zo.zlib_compress(data, level)
    assert zs[:2].hex() == {"none": "7801", "fast": "785e", "default": "789c", "best": "78da"}[level]
    if max(data, default=0) < 115:  # the signed-rem quirk cannot trigger
        assert zlib.decompress(zs) == data and ad == zlib.adler32(data)
    out, ad2 = zo.zlib_decompress(zs)
    assert out == data and ad2 == ad


def test_zlib_decompress_errors():
    _, zs = zo.zlib_compress(b"hello hello hello", "default")
    bad = bytearray(zs); bad[-1] ^= 1
    with pytest.raises(zo.OracleError) as e:
        zo.zlib_decompress(bytes(bad))
    assert e.value.status == zo.ERR_CHECKSUM and "Checksum mismatch, expected" in e.value.message
    with pytest.raises(zo.OracleError) as e:
        zo.zlib_decompress(b"\x78\x9c\x03")
    assert e.value.message == "Corrupted data stream"
    with pytest.raises(zo.OracleError) as e:  # CM = 7
        zo.zlib_decompress(bytes([0x77, 31 - (0x7700 % 31)]) + zs[2:])
    assert e.value.message == "Unknown compression method (7)"
    with pytest.raises(zo.OracleError) as e:  # FDICT
        hdr = 0x7800 | 0x20
        zo.zlib_decompress(bytes([0x78, (hdr + 31 - hdr % 31) & 0xFF]) + zs[2:])
    assert e.value.message == "Preset dictionary unsupported"


CORRUPT = [
    b"",                                # no header bits
    b"\x07",                            # BTYPE 3
    b"\x01\x01\x00\x00\xff",            # stored LEN/NLEN mismatch
    b"\x01\x05\x00\xfa\xff\x01",        # stored block truncated
    b"\x03",                            # fixed block, missing symbols
    b"\x63\x00",                        # fixed: not final and input ends
    bytes.fromhex("4b040000"),          # trailing data after final block is ignored -> ok
]


def test_inflate_corrupt_streams():
    for s in CORRUPT[:-1]:
        with pytest.raises(zo.OracleError) as e:
            zo.inflate(s)
        assert e.value.message == "Corrupted data stream", s
    assert zo.inflate(CORRUPT[-1]) == b"a"
    # a distance reaching before the start of the output: fixed block "length 3 dist 1" at pos 0
    bits = "1" + "10" + "0000001" + "00000"  # BFINAL, BTYPE=01 (lsb first), sym 257, dist sym 0
    v = int(bits[::-1], 2).to_bytes(3, "little")
    with pytest.raises(zo.OracleError):
        zo.inflate(v)


def test_codelen_freqs_switch_changes_only_block_choice():
    data = _corpus()["text"] * 3
    as_written = zo.deflate(data, "default")
    zo.set_keep_codelen_freqs(False)
    try:
        fixed = zo.deflate(data, "default")
    finally:
        zo.set_keep_codelen_freqs(True)
    assert zo.inflate(as_written) == data and zo.inflate(fixed) == data
    assert len(fixed) <= len(as_written)
