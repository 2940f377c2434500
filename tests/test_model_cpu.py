"""The encoder's algorithm, exercised on the CPU through the host model (same building blocks as the
kernel): streams must inflate through the oracle and zlib, and stay within the stated ratio tolerance
of the reference (restated) encoder per level."""
import random
import zlib

import pytest

from oracle import zipc_oracle as zo
from zipc_b200 import synth
import encoder_model as model

RATIO_TOLERANCE = 1.05  # ratio_new <= 1.05 * ratio_ref(per-block reset of code length freqs)


def _corpus():
    rnd = random.Random(5)
    t = synth.text_v1(3, 100000).tobytes()
    return {
        "text": synth.text_v1(1007, 133120).tobytes(),
        "random": rnd.randbytes(70000),
        "mixed": t + rnd.randbytes(50000) + t,
        "zeros": bytes(150000),
        "runs": b"".join(bytes([rnd.randrange(256)]) * rnd.randrange(1, 600) for _ in range(500)),
        "tiny": b"abc", "empty": b"", "one": b"a", "hello": b"hellohello",
    }


@pytest.mark.parametrize("level", ["fast", "default", "best"])
def test_model_roundtrip_and_ratio(level):
    zo.set_keep_codelen_freqs(False)
    try:
        for name, data in _corpus().items():
            cs = model.deflate(data, level)
            assert zo.inflate(cs) == data, name
            assert zlib.decompress(cs, -15) == data, name
            ref = zo.deflate(data, level)
            assert len(cs) <= RATIO_TOLERANCE * len(ref) + 8, (name, len(cs), len(ref))
    finally:
        zo.set_keep_codelen_freqs(True)


def test_model_small_kats_match_reference_bytes():
    # on these the parse and block choice coincide with the reference's (SURVEY.md 8c KATs)
    assert model.deflate(b"", "default").hex() == "0300"
    assert model.deflate(b"a", "default").hex() == "4b0400"
    assert model.deflate(b"hellohello", "default").hex() == "cb48cdc9c9071300"
