"""GPU parity: batched deflate.  The reference pins no deflate bytes, so the bar is:
(1) every stream inflates bit-exactly through the oracle's inflate (stand-in for Zipc_deflate.inflate) and
zlib, (2) checksums of the input are bit-exact, (3) compressed size within RATIO_TOLERANCE of the oracle
per level, (4) the kernel's bytes equal the host model's (same algorithm run serially)."""
import io
import random
import zipfile
import zlib

import numpy as np
import pytest

from oracle import zipc_oracle as zo
import encoder_model as model
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

pytestmark = pytest.mark.gpu
RATIO_TOLERANCE = 1.05
LEVELS = ("none", "fast", "default", "best")


@pytest.fixture(scope="module")
def ctx():
    c = zd.Context(0)
    zd.set_default_context(c)
    yield c
    zd.set_default_context(None)
    c.close()


def test_deflate_trip(ctx):  # test/test.ml:28-43
    strings = [b"", b"a", b"hellohello", b"abcdefghijklmnopqrstuvwxyzzyxwvutsrqponmlkjihgfedcba",
               bytes((i + 1) % 255 for i in range(256))]
    for level in LEVELS:
        for s in strings:
            cs = zd.deflate(s, level=level).get_ok()
            assert zo.inflate(cs) == s and zlib.decompress(cs, -15) == s
            assert zd.inflate(cs).get_ok() == s  # and back through our own decoder
    assert zd.deflate(b"", level="none").get_ok().hex() == "010000ffff"
    assert zd.deflate(b"").get_ok().hex() == "0300"
    assert zd.deflate(b"a").get_ok().hex() == "4b0400"
    assert zd.deflate(b"hellohello").get_ok().hex() == "cb48cdc9c9071300"


def test_decompression_size_limits_on_our_stream(ctx):  # test/test.ml:45-55
    src = b"Keep it to the limits."
    csrc = zd.deflate(src).get_ok()
    assert zo.inflate(csrc, len(src)) == src and zo.inflate(csrc, len(src) + 1) == src
    with pytest.raises(zo.OracleError):
        zo.inflate(csrc, len(src) - 1)


def _corpus():
    rnd = random.Random(5)
    t = synth.text_v1(3, 100000).tobytes()
    return {
        "text": synth.text_v1(1007, 133120).tobytes(),
        "text_big": synth.text_v1(11, 700001).tobytes(),
        "random": rnd.randbytes(70000),
        "mixed": t + rnd.randbytes(50000) + t,
        "zeros": bytes(150000),
        "runs": b"".join(bytes([rnd.randrange(256)]) * rnd.randrange(1, 600) for _ in range(500)),
        "tiny": b"abc", "edge2047": t[:2047], "edge2048": t[:2048], "edge2049": t[:2049],
        "edge61440": t[:61440], "edge61441": t[:61441], "blk": t[:65536] + t[:65536],
    }


@pytest.mark.parametrize("level", ["fast", "default", "best"])
def test_corpus_roundtrip_ratio_and_model_equality(ctx, level, monkeypatch):
    corpus = _corpus()
    names = list(corpus)
    # (the host model is the one-CTA encoder: no member of this thin batch is to be split into primed segments here)
    monkeypatch.setenv("ZIPC_B200_SPLIT_MIN", str(1 << 40))
    res = ctx.deflate_batch([corpus[k] for k in names], level, _lib.CK_CRC32)
    monkeypatch.delenv("ZIPC_B200_SPLIT_MIN")
    # the same batch as a caller gets it: "text_big" (700 KB in a batch that does not fill the GPU) goes over several CTAs
    for k, (st, cs, crc), (_, one, _) in zip(names, ctx.deflate_batch([corpus[k] for k in names], level, _lib.CK_CRC32), res):
        assert st == 0 and crc == zo.crc32(corpus[k]) and zlib.decompress(cs.tobytes(), -15) == corpus[k], k
        assert len(cs) <= len(one) * 1.002 + 8 * 16, k
        assert (len(cs) != len(one)) == (len(corpus[k]) >= (320 << 10)), k
    zo.set_keep_codelen_freqs(False)
    try:
        for k, (st, cs, crc) in zip(names, res):
            data = corpus[k]
            cs = cs.tobytes()
            assert st == 0, k
            assert crc == zo.crc32(data), k
            assert zo.inflate(cs) == data, k
            assert zlib.decompress(cs, -15) == data, k
            assert len(cs) <= RATIO_TOLERANCE * len(zo.deflate(data, level)) + 8, k
            assert cs == model.deflate(data, level), k
    finally:
        zo.set_keep_codelen_freqs(True)


def test_level_none_matches_reference_bytes(ctx):
    """`None: stored blocks of 65,534 source bytes (zipc_deflate.ml:747-750, :1106-1116) -- byte-identical to the oracle"""
    for data in (b"", b"x", synth.text_v1(2, 65534).tobytes(), synth.text_v1(2, 65535).tobytes(), synth.text_v1(2, 65536).tobytes(),
                 synth.rand_v1(3, 200000).tobytes()):
        cs = zd.deflate(data, level="none").get_ok()
        assert cs == zo.deflate(data, "none")
        assert zo.inflate(cs) == data and zlib.decompress(cs, -15) == data
        assert len(cs) == len(data) + 5 * max(1, -(-len(data) // 65534))


def _quirk_corpus():
    """inputs on which the reference's signed Int32.rem Adler-32 differs from RFC 1950 (mean byte >= 115 per chunk), so
    that the per-block fold and the re-packed state between blocks matter (zipc_deflate.ml:175-198, :1081-1086)"""
    t = np.frombuffer(synth.text_v1(21, 300000).tobytes(), dtype=np.uint8)
    rnd = synth.rand_v1(22, 250001).tobytes()
    return {
        "text_bit7": (t | 0x80).tobytes(),
        "rand": rnd,
        "ff_runs": b"\xff" * 200000 + rnd[:5000] + b"\xff" * 70000,
        "mix": (t[:90000] | 0x80).tobytes() + rnd[:70000] + bytes(range(256)) * 300 + b"\xfe" * 65534 + t[:50000].tobytes(),
        "ramp": bytes(range(256)) * 1000,
        "short_ff": b"\xff" * 5552,
        "empty": b"",
    }


@pytest.mark.parametrize("level", LEVELS)
def test_zlib_compress_ref_compat_round_trips_through_the_reference(ctx, level):
    """Z-2 / D-1: in REF_COMPAT mode the trailer (and adler_32_and_deflate's value) is the Adler-32 folded over the
    encoder's own blocks, i.e. exactly what the reference's zlib_decompress recomputes from the stream."""
    corpus = _quirk_corpus()
    names = list(corpus)
    res = ctx.zlib_compress_batch([corpus[k] for k in names], level, _lib.ADLER_REF_COMPAT)
    quirk_seen = 0
    for k, (st, zs, ad) in zip(names, res):
        data, zs = corpus[k], zs.tobytes()
        assert st == 0, k
        out, found = zo.zlib_decompress(zs)              # raises on "Checksum mismatch"
        assert out == data and found == ad, k
        assert int.from_bytes(zs[-4:], "big") == ad, k
        assert zd.zlib_decompress(zs).get_ok() == (data, ad), k   # and our own decoder agrees
        quirk_seen += ad != zlib.adler32(data)
    assert quirk_seen >= 3  # the corpus does exercise the quirk
    # adler_32_and_deflate returns the same per-block value
    res = ctx.deflate_batch([corpus[k] for k in names], level, _lib.CK_ADLER32, _lib.ADLER_REF_COMPAT)
    for k, (st, ds, ad) in zip(names, res):
        assert st == 0 and zo.inflate_and_adler_32(ds.tobytes()) == (corpus[k], ad), k
    # RFC 1950 mode is the standard checksum: zlib reads the streams
    res = ctx.zlib_compress_batch([corpus[k] for k in names], level, _lib.ADLER_RFC1950)
    for k, (st, zs, ad) in zip(names, res):
        assert st == 0 and ad == zlib.adler32(corpus[k]) and zlib.decompress(zs.tobytes()) == corpus[k], k


def test_zlib_compress_of_large_payloads_is_split_and_still_the_references_adler(ctx, monkeypatch):
    """A large payload is compressed as primed segments, one CTA each, under the zlib framing too: the trailer is the Adler-32
    folded over the blocks of all segments in stream order -- what the reference's zlib_decompress recomputes from the stream
    (on data where its signed remainder differs from RFC 1950) -- and the call takes a fraction of the time one CTA needs
    (~35 ms per 3 MiB)."""
    import time
    t = np.frombuffer(synth.text_v1(23, 3 << 20).tobytes(), dtype=np.uint8)
    big = (t | 0x80).tobytes()[:-777] + b"\xff" * 100_000 + synth.rand_v1(24, 50_001).tobytes()
    for level in ("fast", "default"):
        (st, zs, ad), = ctx.zlib_compress_batch([big], level, _lib.ADLER_REF_COMPAT)
        zs = zs.tobytes()
        assert st == 0 and int.from_bytes(zs[-4:], "big") == ad
        out, found = zo.zlib_decompress(zs)              # raises on "Checksum mismatch"
        assert out == big and found == ad and ad != zlib.adler32(big)
        assert zd.zlib_decompress(zs).get_ok() == (big, ad)
        (st, ds, ad2), = ctx.deflate_batch([big], level, _lib.CK_ADLER32, _lib.ADLER_REF_COMPAT)
        assert st == 0 and ad2 == ad and ds.tobytes() == zs[2:-4]
        (st, zs2, ad3), = ctx.zlib_compress_batch([big], level, _lib.ADLER_RFC1950)
        assert st == 0 and ad3 == zlib.adler32(big) and zlib.decompress(zs2.tobytes()) == big
    def timed():
        t0 = time.perf_counter()
        ctx.zlib_compress_batch([big], "default", _lib.ADLER_REF_COMPAT)
        return time.perf_counter() - t0
    split = min(timed() for _ in range(3))
    monkeypatch.setenv("ZIPC_B200_SPLIT_MIN", str(1 << 40))   # the same call left to one CTA
    timed()
    one = min(timed() for _ in range(2))
    monkeypatch.delenv("ZIPC_B200_SPLIT_MIN")
    assert split * 3 < one, (split, one)


def test_fixture_redeflate_recode(ctx, zip_docs):  # test/test.ml:58-118 with the GPU codec in the loop
    ms = [m for m in zo.zip_decode(zip_docs)]
    out = []
    for m in ms:
        if m.is_dir:
            out.append(m)
            continue
        s = zd.inflate_and_crc_32(zip_docs, start=m.start, len=m.compressed_size, decompressed_size=m.decompressed_size).get_ok()
        assert s[1] == m.crc32
        crc, cs = zd.crc_32_and_deflate(s[0], level="best").get_ok()
        assert crc == m.crc32
        out.append(zo.member_make(m.path, mode=m.mode, mtime=m.mtime, compression=8, compressed_bytes=cs,
                                  decompressed_size=len(s[0]), crc32=crc))
    recoded = zo.zip_encode(out)
    back = {m.path: m for m in zo.zip_decode(recoded)}
    assert back[b"zip-docs/rfc1951.txt"].crc32 == 0xFB4F3400 and back[b"zip-docs/APPNOTE.TXT"].crc32 == 0x39B029C4
    for m in back.values():
        if not m.is_dir:
            zo.file_to_binary_string(m)  # inflate + CRC check by the oracle
    assert zipfile.ZipFile(io.BytesIO(recoded)).testzip() is None


def test_zlib_compress(ctx):
    text = synth.text_v1(6, 100000).tobytes()
    for level, hdr in (("none", "7801"), ("fast", "785e"), ("default", "789c"), ("best", "78da")):
        ad, zs = zd.zlib_compress(text, level=level).get_ok()
        assert zs[:2].hex() == hdr and ad == zlib.adler32(text) == zo.adler32(text)
        assert zlib.decompress(zs) == text
        assert zo.zlib_decompress(zs) == (text, ad)
        assert zd.zlib_decompress(zs).get_ok() == (text, ad)
    rnd = synth.rand_v1(8, 100000).tobytes()
    ad, zs = zd.zlib_compress(rnd, level="default", adler_mode=_lib.ADLER_RFC1950).get_ok()
    assert zlib.decompress(zs) == rnd and ad == zlib.adler32(rnd)
    crc, ds = zd.adler_32_and_deflate(text).get_ok()
    assert crc == zlib.adler32(text) and zo.inflate(ds) == text


def test_many_members_ragged(ctx):
    """C4 shape, scaled down: 400 members of 4-256 KiB (10 % incompressible), three levels."""
    sizes = synth.member_sizes(400, seed=9)
    datas = [(synth.rand_v1(2000 + i, int(n)) if i % 10 == 0 else synth.text_v1(2000 + i, int(n))) for i, n in enumerate(sizes)]
    total = sum(d.size for d in datas)
    zo.set_keep_codelen_freqs(False)
    try:
        for level in ("fast", "default", "best"):
            res = ctx.deflate_batch(datas, level, _lib.CK_CRC32)
            csum = 0
            for d, (st, cs, crc) in zip(datas, res):
                assert st == 0 and crc == zlib.crc32(d.tobytes())
                assert zlib.decompress(cs.tobytes(), -15) == d.tobytes()
                csum += cs.size
            ref = sum(len(zo.deflate(d.tobytes(), level)) for d in datas[:40])
            ours = sum(r[1].size for r in res[:40])
            assert ours <= RATIO_TOLERANCE * ref
            # and our own decoder reads them back bit-exactly
            back = ctx.inflate_batch([r[1] for r in res], [d.size for d in datas], _lib.CK_CRC32)
            for d, r, (st, out, crc) in zip(datas, res, back):
                assert st == 0 and crc == r[2] and (out == d).all()
    finally:
        zo.set_keep_codelen_freqs(True)


def test_segmented_single_stream(ctx):
    """C1 / C5 shape, scaled down: one stream compressed as independent segments; the concatenation is ONE valid
    RFC 1951 stream for the reference's inflate (oracle) and zlib, and the index lets us inflate it in parallel."""
    data = synth.text_v1(1, (6 << 20) + 12345)
    b = data.tobytes()
    for seg in (64 << 10, 256 << 10, 1 << 20):
        stream, index, crc = ctx.deflate_segmented(data, "default", seg)
        assert crc == zlib.crc32(b)
        assert index.shape[0] == -(-data.size // seg) + 1 and int(index[-1, 0]) == stream.size and int(index[-1, 1]) == data.size
        s = stream.tobytes()
        assert zlib.decompress(s, -15) == b
        if seg == 256 << 10:
            assert zo.inflate(s) == b                       # Zipc_deflate.inflate stand-in
            out, ocrc = zo.inflate_and_crc_32(s, len(b))
            assert ocrc == crc
        st, out, crc2 = ctx.inflate_segmented(stream, index)
        assert st == 0 and crc2 == crc and (out == data).all()
        # serial decode of the same stream through the ordinary entry point
        if seg == 1 << 20:
            r = zd.inflate_and_crc_32(s, decompressed_size=len(b))
            assert r.get_ok() == (b, crc)
        ref = len(zlib.compress(b, 6))
        assert stream.size <= 1.10 * ref
    # pieces for several GPUs: non-final slices concatenate to one stream
    half = data.size // 2 // 512 * 512
    a, ia, ca = ctx.deflate_segmented(data[:half], "fast", 128 << 10, last_piece=False)
    c, ic, cc = ctx.deflate_segmented(data[half:], "fast", 128 << 10, last_piece=True)
    joined = a.tobytes() + c.tobytes()
    assert zlib.decompress(joined, -15) == b and zo.inflate(joined) == b
    assert _lib.lib().zipc_b200_crc32_combine(ca, cc, data.size - half) == zlib.crc32(b)
    # edge cases
    for n in (0, 1, 4096, 65536, 65537):
        stream, index, crc = ctx.deflate_segmented(data[:n], "default", 64 << 10)
        assert zlib.decompress(stream.tobytes(), -15) == b[:n] and crc == zlib.crc32(b[:n])
        st, out, crc2 = ctx.inflate_segmented(stream, index)
        assert st == 0 and out.tobytes() == b[:n]
    # a damaged segment is reported, the call itself survives
    stream, index, crc = ctx.deflate_segmented(data, "default", 256 << 10)
    bad = stream.copy(); bad[int(index[3, 0]) + 100] ^= 0xFF
    st, out, _ = ctx.inflate_segmented(bad, index)
    assert st != 0 or out.tobytes() != b


@pytest.mark.parametrize("level", ["fast", "default", "best"])
def test_primed_segments_lose_no_ratio(ctx, level):
    """zipc_b200_deflate_primed: one stream compressed by one CTA per segment, every segment primed with the 32 KiB of input
    before it.  The result is ONE ordinary RFC 1951 stream (the oracle's and zlib's inflate read it), as small as the same
    input compressed in one piece (5 bytes per segment for the byte-aligning stored blocks), and smaller than with window
    resets; the GPU's own foreign-stream decoder reads it back."""
    n = (3 << 20) + 12345 if level != "best" else (1 << 20) + 777
    s = synth.text_v1(77, n).tobytes()
    whole = bytes(zd.deflate(s, level=level).get_ok())
    for seg in (64 << 10, (96 << 10) + 1000):   # (the second: segment starts that are no multiples of the tile)
        primed, index, crc = ctx.deflate_segmented(s, level, seg, primed=True)
        reset, _, _ = ctx.deflate_segmented(s, level, seg, primed=False)
        nseg = index.shape[0] - 1
        assert crc == zlib.crc32(s)
        assert zlib.decompress(bytes(primed), -15) == s
        assert zo.inflate(bytes(primed)) == s
        assert len(primed) <= len(whole) * 1.002 + 8 * nseg, (len(primed), len(whole))
        assert len(primed) < len(reset)
        out = zd.inflate(bytes(primed), decompressed_size=len(s)).get_ok()
        assert bytes(out) == s
    # a non-final piece (a slice of a stream spread over several GPUs) keeps BFINAL off
    piece, _, _ = ctx.deflate_segmented(s[:300_000], level, 64 << 10, last_piece=False, primed=True)
    d = zlib.decompressobj(-15)
    assert d.decompress(bytes(piece)) == s[:300_000] and not d.eof


def test_large_members_are_split_over_ctas(ctx, monkeypatch):
    """A member of 2 MiB or more is compressed as primed 256 KiB segments, one CTA each (Zipc.File.deflate_of_binary_string of
    one big payload must not be left to a single SM): still ONE ordinary stream per member, CRC-32 of the whole member,
    5 bytes per segment larger than the one-CTA result at most, and much faster."""
    import time
    big = [synth.text_v1(31, (5 << 20) + 4321).tobytes(), synth.rand_v1(32, (2 << 20) + 17).tobytes(), synth.text_v1(33, 2 << 20).tobytes()]
    small = [synth.text_v1(40 + i, 50_000 + 1000 * i).tobytes() for i in range(6)]
    items = [small[0], big[0], small[1], small[2], big[1], big[2], small[3], small[4], small[5]]
    for level in ("fast", "default"):
        res = ctx.deflate_batch(items, level, _lib.CK_CRC32)
        for d, (st, cs, ck) in zip(items, res):
            assert st == 0 and ck == zlib.crc32(d)
            assert zlib.decompress(bytes(cs), -15) == d
            if len(d) >= (2 << 20):
                assert zo.inflate(bytes(cs)) == d
        # against the same member compressed by one CTA (below the threshold: the first 2 MiB - 1 of it, scaled)
        d = big[0]
        monkeypatch.setenv("ZIPC_B200_SPLIT_MIN", str(1 << 40))   # (one CTA for the comparison)
        one = ctx.deflate_batch([d[:(2 << 20) - 1]], level, 0)[0][1]
        monkeypatch.delenv("ZIPC_B200_SPLIT_MIN")
        split = res[1][1]
        assert len(split) / len(d) <= len(one) / ((2 << 20) - 1) * 1.01
        # a thin batch splits from 256 KiB on: the same bytes within 5 bytes per segment of the one-CTA stream
        again = ctx.deflate_batch([d[:(2 << 20) - 1]], level, _lib.CK_CRC32)[0]
        assert again[0] == 0 and again[2] == zlib.crc32(d[:(2 << 20) - 1]) and zlib.decompress(bytes(again[1]), -15) == d[:(2 << 20) - 1]
        assert len(one) < len(again[1]) <= len(one) * 1.002 + 8 * 32
    # one 64 MiB payload: seconds on one SM, milliseconds on all of them
    huge = synth.text_v1(5, 64 << 20)
    ctx.deflate_batch([huge], "default", _lib.CK_CRC32)
    t0 = time.perf_counter()
    st, cs, ck = ctx.deflate_batch([huge], "default", _lib.CK_CRC32)[0]
    dt = time.perf_counter() - t0
    assert st == 0 and ck == zlib.crc32(huge) and zlib.decompress(bytes(cs), -15) == huge.tobytes()
    assert dt < 0.25, dt   # (one CTA: ~0.75 s)
    # back through the many-warp decoder -- not the one-warp fallback: this very stream holds an accidental block header
    # inside a block, which breaks the chain of chunks once; the decoder must re-cut from there and go on in parallel
    par0, fb0 = ctx.parallel_streams
    assert zd.inflate(bytes(cs), decompressed_size=huge.size).get_ok() == huge.tobytes()
    par1, fb1 = ctx.parallel_streams
    assert par1 > par0 and fb1 == fb0
