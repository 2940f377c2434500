"""GPU parity: batched inflate vs the oracle -- output bytes, checksums and error statuses."""
import io
import random
import zipfile
import zlib

import numpy as np
import pytest

from oracle import zipc_oracle as zo
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zd.Context(0)
    zd.set_default_context(c)
    yield c
    zd.set_default_context(None)
    c.close()


def _oracle_inflate(s, dsize, crc_op):
    try:
        out, crc = zo.inflate_and_crc(bytes(s), dsize, {0: zo.CRC_NOP, 1: zo.CRC_ADLER32, 2: zo.CRC_CRC32}[crc_op])
        return 0, out, crc
    except zo.OracleError as e:
        return e.status, b"", 0


def _check_batch(ctx, streams, dsizes, crc_op=_lib.CK_CRC32):
    got = ctx.inflate_batch(streams, dsizes, crc_op)
    for i, (st, out, ck) in enumerate(got):
        est, eout, ecrc = _oracle_inflate(streams[i], dsizes[i] if dsizes else None, crc_op)
        assert st == est, (i, st, est)
        if st == 0:
            assert out.tobytes() == eout, i
            assert ck == ecrc, i


def test_deflate_trip_strings(ctx):  # test/test.ml:28-43 (decode side: oracle-made streams)
    strings = [b"", b"a", b"hellohello", b"abcdefghijklmnopqrstuvwxyzzyxwvutsrqponmlkjihgfedcba",
               bytes((i + 1) % 255 for i in range(256))]
    streams = [zo.deflate(s, lvl) for s in strings for lvl in ("none", "fast", "default", "best")]
    _check_batch(ctx, streams, None)
    _check_batch(ctx, streams, [len(s) for s in strings for _ in range(4)])
    for s in strings:
        assert zd.inflate(zo.deflate(s, "default")).get_ok() == s


def test_decompression_size_limits(ctx):  # test/test.ml:45-55
    src = b"Keep it to the limits."
    csrc = zo.deflate(src)
    assert zd.inflate(csrc).get_ok() == src
    assert zd.inflate(csrc, decompressed_size=len(src)).get_ok() == src
    assert zd.inflate(csrc, decompressed_size=len(src) + 1).get_ok() == src
    r = zd.inflate(csrc, decompressed_size=len(src) - 1)
    assert r.is_error() and r.message == "Expected decompression size exceeded"


def test_fixture_members(ctx, zip_docs):  # test/test.ml:57-118, decode side
    ms = [m for m in zo.zip_decode(zip_docs) if not m.is_dir]
    streams = [zip_docs[m.start:m.start + m.compressed_size] for m in ms]
    got = ctx.inflate_batch(streams, [m.decompressed_size for m in ms], _lib.CK_CRC32)
    for m, (st, out, ck) in zip(ms, got):
        assert st == 0 and ck == m.crc32 and out.tobytes() == zo.file_to_binary_string(m)
    assert {m.crc32 for m in ms} == {0xFB4F3400, 0x39B029C4}
    r = zd.inflate_and_crc_32(zip_docs, start=ms[0].start, len=ms[0].compressed_size)
    assert r.get_ok()[1] == ms[0].crc32


def _corpus():
    rnd = random.Random(11)
    text = synth.text_v1(3, 300000).tobytes()
    return {
        "text": text,
        "random": rnd.randbytes(150000),
        "zeros": bytes(200000),
        "runs": b"".join(bytes([rnd.randrange(256)]) * rnd.randrange(1, 600) for _ in range(800)),
        "mixed": text[:70000] + rnd.randbytes(66000) + text[:70000] + bytes(5000),
        "short": b"abc",
        "period3": b"abc" * 30000,
    }


def test_zlib_made_streams_all_block_kinds(ctx):
    streams, sizes = [], []
    for name, data in _corpus().items():
        for lvl in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strategy)
                streams.append(c.compress(data) + c.flush())
                sizes.append(len(data))
    _check_batch(ctx, streams, sizes, _lib.CK_CRC32)
    _check_batch(ctx, streams, None, _lib.CK_ADLER32)
    _check_batch(ctx, streams, None, _lib.CK_NONE)


def test_oracle_made_streams(ctx):
    streams, sizes = [], []
    for name, data in _corpus().items():
        for lvl in ("none", "fast", "default", "best"):
            streams.append(zo.deflate(data, lvl))
            sizes.append(len(data))
    _check_batch(ctx, streams, sizes)


def test_extreme_code_shapes(ctx):
    """Streams chosen for the speculative decoder: 15-bit and 1-bit Huffman codes (geometric byte
    distributions), runs of one byte (1-bit literal / 258-byte matches at distance 1, more than 32 tokens per
    256-bit window), far distances with long distance codes, and many short blocks with changing tables."""
    rng = np.random.default_rng(11)
    datas = []
    # geometric distributions: P(b) ~ 2^-k -> code lengths 1..15
    for scale in (0.35, 0.7, 1.4, 3.0):
        datas.append(np.minimum(rng.exponential(scale, 300000), 255).astype(np.uint8).tobytes())
    datas.append(bytes(300000))                                            # one symbol
    datas.append((b"\x00" * 1000 + b"\x01") * 300)                      # near-degenerate code
    datas.append(bytes(rng.integers(0, 2, 200000, dtype=np.uint8)))        # two symbols: 1-bit literals
    datas.append(bytes(rng.integers(0, 4, 200000, dtype=np.uint8)))
    far = rng.integers(0, 256, 33000, dtype=np.uint8).tobytes()
    datas.append(far + far[:5000] + far[20000:29000] + far[:33000])        # distances up to 32768
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(5000)]
    datas.append(b" ".join(words[int(i)] for i in rng.zipf(1.3, 60000) % 5000))
    streams, sizes = [], []
    for d in datas:
        for lvl, strategy in ((9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_HUFFMAN_ONLY),
                              (6, zlib.Z_RLE), (6, zlib.Z_FILTERED), (6, zlib.Z_FIXED)):
            c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strategy)
            streams.append(c.compress(d) + c.flush()); sizes.append(len(d))
        for lvl in ("fast", "best"):
            streams.append(zo.deflate(d, lvl)); sizes.append(len(d))
        # many short blocks, tables change every 1000 bytes
        c = zlib.compressobj(6, zlib.DEFLATED, -15, 1)
        out = b"".join(c.compress(d[i:i + 1000]) + c.flush(zlib.Z_FULL_FLUSH if (i // 1000) % 2 else zlib.Z_SYNC_FLUSH)
                       for i in range(0, min(len(d), 60000), 1000)) + c.flush()
        streams.append(out); sizes.append(min(len(d), 60000))
    got = ctx.inflate_batch(streams, sizes, _lib.CK_CRC32)
    k = 0
    for d in datas:
        for _ in range(9):
            st, out, ck = got[k]
            exp = d[:sizes[k]]
            assert st == 0, (k, st)
            assert out.tobytes() == exp, k
            assert ck == zlib.crc32(exp), k
            k += 1
    # unknown sizes (count pass) and a too-small limit on every stream
    got = ctx.inflate_batch(streams, None, _lib.CK_ADLER32)
    for k, (st, out, ck) in enumerate(got):
        assert st == 0 and len(out) == sizes[k], k
    got = ctx.inflate_batch(streams, [max(0, n - 1) for n in sizes], _lib.CK_NONE)
    assert all(st == _lib.ERR_SIZE_EXCEEDED for st, _, _ in got)


def test_multi_member_sync_flush_streams(ctx):
    # many small blocks incl. empty stored blocks (Z_SYNC_FLUSH) and block type changes
    data = synth.text_v1(9, 200000).tobytes()
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    out = b""
    for i in range(0, len(data), 7777):
        out += c.compress(data[i:i + 7777]) + c.flush(zlib.Z_SYNC_FLUSH if i % 3 else zlib.Z_FULL_FLUSH)
    out += c.flush()
    _check_batch(ctx, [out], [len(data)])


def _mutations(stream: bytes, rnd, count):
    out = []
    for _ in range(count):
        b = bytearray(stream)
        kind = rnd.randrange(4)
        if kind == 0 and len(b) > 1:
            del b[rnd.randrange(1, len(b)):]              # truncate
        elif kind == 1:
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)   # bit flips
        elif kind == 2:
            p = rnd.randrange(len(b)); b[p:p + rnd.randrange(1, 9)] = rnd.randbytes(rnd.randrange(1, 9))
        else:
            b[:rnd.randrange(1, 6)] = rnd.randbytes(rnd.randrange(1, 6))   # header damage
        out.append(bytes(b))
    return out


def test_corrupt_streams_status_parity(ctx):
    rnd = random.Random(2024)
    text = synth.text_v1(4, 40000).tobytes()
    bases = [zo.deflate(text, "default"), zo.deflate(text[:300], "fast"), zlib.compress(text, 1)[2:-4],
             zo.deflate(rnd.randbytes(3000), "default"), zo.deflate(b"hellohello", "default")]
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
    bases.append(c.compress(text[:5000]) + c.flush())
    streams = [b"", b"\x07", b"\x01\x01\x00\x00\xff", b"\x01\x05\x00\xfa\xff\x01", b"\x03", b"\x63\x00",
               bytes.fromhex("4b040000"), b"\x05", b"\x04\x00", b"\xff" * 40, b"\x00" * 40]
    for b in bases:
        streams += _mutations(b, rnd, 120)
    sizes = [None] * len(streams)
    _check_batch(ctx, streams, sizes)
    # with a decompressed_size that is sometimes too small: both error kinds, in the reference's order
    sizes2 = [rnd.choice([0, 1, 100, 299, 300, 4999, 5000, 39999, 40000, 50000]) for _ in streams]
    _check_batch(ctx, streams, sizes2)


def test_zlib_decompress(ctx):
    text = synth.text_v1(6, 100000).tobytes()
    ad, zs = zo.zlib_compress(text, "default")
    r = zd.zlib_decompress(zs)
    assert r.get_ok() == (text, ad)
    assert zd.zlib_decompress(zlib.compress(text, 9)).get_ok() == (text, zlib.adler32(text))
    bad = bytearray(zs); bad[-1] ^= 1
    r = zd.zlib_decompress(bytes(bad))
    assert r.is_error() and r.message == "Checksum mismatch, expected %x found %x)" % (ad ^ 1, ad) and r.info == (ad ^ 1, ad)
    assert zd.zlib_decompress(b"\x78\x9c\x03").message == "Corrupted data stream"
    assert zd.zlib_decompress(bytes([0x77, 31 - (0x7700 % 31)]) + zs[2:]).message == "Unknown compression method (7)"
    hdr = 0x7800 | 0x20
    assert zd.zlib_decompress(bytes([0x78, (hdr + 31 - hdr % 31) & 0xFF]) + zs[2:]).message == "Preset dictionary unsupported"
    assert zd.zlib_decompress(bytes([0x88, 31 - (0x8800 % 31)]) + zs[2:]).message == "Window size too large"
    for s in (zs, bytes(bad), b"\x78\x9c\x03", zs[:20]):
        try:
            zo.zlib_decompress(s); ok = True
        except zo.OracleError:
            ok = False
        assert zd.zlib_decompress(s).is_ok() == ok


def test_many_members_ragged(ctx):
    """C3 shape, scaled down: 600 members of 4-256 KiB text (+ some incompressible), zlib- and
    oracle-made, decoded in one batch; every output and CRC bit-exact."""
    sizes = synth.member_sizes(600, seed=7)
    streams, dsz, crcs = [], [], []
    for i, n in enumerate(sizes):
        n = int(n)
        data = (synth.rand_v1(1000 + i, n) if i % 10 == 0 else synth.text_v1(1000 + i, n)).tobytes()
        if i % 3 == 0 and n < 60000:
            cs = zo.deflate(data, "default")
        else:
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            cs = c.compress(data) + c.flush()
        streams.append(cs); dsz.append(n); crcs.append(zlib.crc32(data))
    got = ctx.inflate_batch(streams, dsz, _lib.CK_CRC32)
    for i, (st, out, ck) in enumerate(got):
        assert st == 0 and ck == crcs[i] and out.size == dsz[i], i
        assert zlib.crc32(out.tobytes()) == crcs[i], i
