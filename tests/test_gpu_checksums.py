"""GPU parity: CRC-32 / Adler-32 kernels vs the oracle (bit-exact), through the C ABI."""
import zlib

import numpy as np
import pytest

from oracle import zipc_oracle as zo
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

pytestmark = pytest.mark.gpu

FOX = b"The quick brown fox jumps over the lazy dog"


@pytest.fixture(scope="module")
def ctx():
    c = zd.Context(0)
    zd.set_default_context(c)
    yield c
    zd.set_default_context(None)
    c.close()


def test_reference_vectors(ctx):  # test/test.ml:16-26
    assert zd.Crc_32.string(b"") == 0
    assert zd.Crc_32.string(FOX) == 0x414FA339
    assert zd.Adler_32.string(b"") == 1
    assert zd.Adler_32.string(FOX) == 0x5BDC0FDA
    assert zd.Crc_32.check(expect=1, found=2).message == "Checksum mismatch, expected 1 found 2)"


SIZES = [1, 2, 3, 4, 5, 15, 16, 17, 63, 64, 65, 127, 511, 512, 513, 1023, 4095, 4096, 4097, 5551, 5552, 5553,
         8191, 65535, 65536, 100003, 262144 + 7, 1 << 20, (1 << 20) + 13, 3 * 5552 * 16, 6 * 1024 * 1024 + 1]


@pytest.mark.parametrize("n", SIZES)
def test_crc32_sizes_and_alignments(ctx, n):
    buf = synth.rand_v1(n, n + 32)
    for shift in (0, 1, 5, 16, 19):
        v = buf[shift:shift + n]
        assert ctx.crc32(v) == zo.crc32(v.tobytes()) == zlib.crc32(v.tobytes())


@pytest.mark.parametrize("n", SIZES)
def test_adler32_sizes_and_alignments(ctx, n):
    buf = synth.rand_v1(n + 1, n + 32)
    for shift in (0, 3, 16):
        v = buf[shift:shift + n]
        b = v.tobytes()
        assert ctx.adler32(v, _lib.ADLER_REF_COMPAT) == zo.adler32(b)       # signed-rem quirk, bit for bit
        assert ctx.adler32(v, _lib.ADLER_RFC1950) == zlib.adler32(b)


def test_adler32_quirk_kats(ctx):  # SURVEY.md 8a row A-n
    for data, as_written, rfc in [(b"\xff" * 5552, 0xF0BD9B8C, 0xF18F9B8C), (b"\xff" * 100000, 0x04D7302C, 0x149A302C),
                                  (bytes(range(256)) * 1000, 0x9142292F, 0xA73B292F), (b"a" * 10 ** 6, 0x15D870F9, 0x15D870F9)]:
        assert ctx.adler32(data, _lib.ADLER_REF_COMPAT) == as_written
        assert ctx.adler32(data, _lib.ADLER_RFC1950) == rfc


def test_adler32_device_fold_adversarial(ctx):
    """The fold of per-chunk sums runs on the device; inputs chosen so that every class of chunk occurs: certain
    no-wrap, certain wrap-to-negative (0xFF runs), low-weight chunks after a negative state (zeros after 0xFF),
    and chunks within 65521 of the 2^31 boundary (ramps)."""
    rng = np.random.default_rng(7)
    pieces = []
    for i in range(400):
        kind = i % 8
        n = int(rng.integers(1, 40000))
        if kind == 0: pieces.append(np.full(n, 0xFF, np.uint8))
        elif kind == 1: pieces.append(np.zeros(n, np.uint8))
        elif kind == 2: pieces.append(rng.integers(0, 256, n, dtype=np.uint8))
        elif kind == 3: pieces.append(np.full(n, int(rng.integers(120, 136)), np.uint8))   # B near 2^31 per chunk
        elif kind == 4: pieces.append(np.full(n, 1, np.uint8))
        elif kind == 5: pieces.append((np.arange(n) % 251).astype(np.uint8))
        elif kind == 6: pieces.append(rng.integers(100, 160, n, dtype=np.uint8))
        else: pieces.append(np.full(n, int(rng.integers(0, 256)), np.uint8))
    data = np.concatenate(pieces)
    for lo, hi in [(0, len(data)), (1, len(data) - 3), (5552 * 7 + 11, 5552 * 900)]:
        v = data[lo:hi]
        b = v.tobytes()
        assert ctx.adler32(v, _lib.ADLER_REF_COMPAT) == zo.adler32(b)
        assert ctx.adler32(v, _lib.ADLER_RFC1950) == zlib.adler32(b)
    for fill in (0x00, 0x80, 0x81, 0x7F, 0xFF):   # 0x80/0x81: every chunk sits at the wrap boundary
        v = np.full(12 * 1024 * 1024 + 5, fill, np.uint8)
        assert ctx.adler32(v, _lib.ADLER_REF_COMPAT) == zo.adler32(v.tobytes())
        assert ctx.adler32(v, _lib.ADLER_RFC1950) == zlib.adler32(v.tobytes())


def test_crc32_batch_ragged(ctx):
    rng = np.random.default_rng(5)
    items = [synth.rand_v1(i, int(s)).tobytes() for i, s in enumerate(rng.integers(0, 70000, size=300))]
    items += [b"", b"a", FOX, bytes(100000)]
    got = ctx.crc32_batch(items)
    assert got == [zo.crc32(x) for x in items]


def test_crc32_large_text_and_random(ctx):
    for data in (synth.text_v1(1, 64 << 20), synth.rand_v1(2, (64 << 20) + 12345)):
        assert ctx.crc32(data) == zlib.crc32(data)


def test_full_size_property_crc_of_crcs(ctx):
    """C2 at full size (1 GiB): compare with zlib, and check the checksum-of-slices identity
    crc(A||B) = combine(crc(A), crc(B), |B|) that the multi-GPU gather relies on."""
    n = 1 << 30
    data = synth.rand_v1(2, n)
    whole = ctx.crc32(data)
    assert whole == zlib.crc32(data)
    L = _lib.lib()
    acc = None
    for k in range(8):
        part = ctx.crc32(data[k * (n // 8):(k + 1) * (n // 8)])
        acc = part if acc is None else L.zipc_b200_crc32_combine(acc, part, n // 8)
    assert acc == whole
    assert ctx.adler32(data, _lib.ADLER_RFC1950) == zlib.adler32(data)
    assert ctx.adler32(data[:200_000_000], _lib.ADLER_REF_COMPAT) == zo.adler32(data[:200_000_000].tobytes())
