"""GPU: the zipc-tool-equivalent harness (tools/zipc_tool.py; reference test/zipc_tool.ml, DEVEL.md:18-30) over the
C ABI: archive testing, recoding with in-memory check, zlib/deflate file codecs, exit codes."""
import io
import os
import sys
import zipfile
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import zipc_tool  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture()
def docs_zip(tmp_path, zip_docs):
    p = tmp_path / "zip-docs.zip"
    p.write_bytes(zip_docs)
    return str(p)


def test_unzip_test_and_list(docs_zip, capsys):
    assert zipc_tool.main(["unzip", "-t", "-v", docs_zip]) == zipc_tool.EXIT_OK
    err = capsys.readouterr().err
    assert err.count("[ OK ]") == 2 and "[----]" in err and "No errors in" in err
    assert zipc_tool.main(["list", docs_zip]) == 0
    names = capsys.readouterr().out.split()
    assert names == sorted(zipfile.ZipFile(docs_zip).namelist())
    assert zipc_tool.main(["unzip", "-t", docs_zip, "no/such/path"]) == zipc_tool.EXIT_PATH


def test_unzip_detects_corruption(tmp_path, zip_docs):
    zf = zipfile.ZipFile(io.BytesIO(zip_docs))
    info = max(zf.infolist(), key=lambda i: i.compress_size)
    b = bytearray(zip_docs)
    b[info.header_offset + 30 + len(info.filename) + 28 + info.compress_size // 2] ^= 0x5A   # inside the payload
    p = tmp_path / "bad.zip"
    p.write_bytes(bytes(b))
    assert zipc_tool.main(["unzip", "-t", str(p)]) == zipc_tool.EXIT_CORRUPTED


@pytest.mark.parametrize("level", ["fast", "default", "best"])
def test_recode_check_and_output(tmp_path, docs_zip, zip_docs, level):
    assert zipc_tool.main(["recode", "--deflate", "--level", level, "-t", docs_zip]) == zipc_tool.EXIT_OK
    out = tmp_path / "re.zip"
    assert zipc_tool.main(["recode", "--deflate", "--level", level, "-o", str(out), docs_zip]) == 0
    a, b = zipfile.ZipFile(io.BytesIO(zip_docs)), zipfile.ZipFile(str(out))
    assert a.namelist() and sorted(a.namelist()) == sorted(b.namelist())
    assert b.testzip() is None                                   # an independent reader accepts the recoded archive
    for n in a.namelist():
        assert a.read(n) == b.read(n)
    assert zipc_tool.main(["recode", "--stored", "-t", docs_zip]) == 0


def test_zip_unzip_roundtrip(tmp_path):
    src = tmp_path / "src"
    (src / "sub").mkdir(parents=True)
    files = {"a.txt": b"hellohello" * 1000, "sub/b.bin": os.urandom(50000), "sub/empty": b""}
    for n, d in files.items():
        (src / n).write_bytes(d)
    out = tmp_path / "o.zip"
    assert zipc_tool.main(["zip", "-o", str(out), "--strip-prefix", str(src), str(src)]) == 0
    z = zipfile.ZipFile(str(out))
    assert z.testzip() is None and {n: z.read(n) for n in z.namelist()} == files
    dst = tmp_path / "dst"
    assert zipc_tool.main(["unzip", "-d", str(dst), str(out)]) == 0
    for n, d in files.items():
        assert (dst / n).read_bytes() == d


def test_file_codecs_and_crc(tmp_path, capsys):
    data = (b"The quick brown fox jumps over the lazy dog. " * 3000)
    f, c, d = tmp_path / "in", tmp_path / "c", tmp_path / "d"
    f.write_bytes(data)
    assert zipc_tool.main(["compress", "--zlib", str(f), str(c)]) == 0
    assert zlib.decompress(c.read_bytes()) == data                # zlib reads our zlib stream
    assert zipc_tool.main(["decompress", "--zlib", str(c), str(d)]) == 0 and d.read_bytes() == data
    c.write_bytes(zlib.compress(data, 9))
    assert zipc_tool.main(["decompress", "--zlib", str(c), str(d)]) == 0 and d.read_bytes() == data
    assert zipc_tool.main(["compress", "--level", "best", str(f), str(c)]) == 0
    assert zlib.decompress(c.read_bytes(), -15) == data
    assert zipc_tool.main(["decompress", str(c), str(d)]) == 0 and d.read_bytes() == data
    capsys.readouterr()
    assert zipc_tool.main(["crc", str(f)]) == 0
    assert int(capsys.readouterr().out.strip(), 16) == zlib.crc32(data)
    assert zipc_tool.main(["crc", "--adler-32", str(f)]) == 0
    assert int(capsys.readouterr().out.strip(), 16) == zlib.adler32(data)
    c.write_bytes(b"\x78\x9c\x00garbage")
    assert zipc_tool.main(["decompress", "--zlib", str(c), str(d)]) == zipc_tool.EXIT_SOME


def test_foreign_archives_and_unsupported_members(tmp_path):
    """Archives written by another implementation (python zipfile): deflate at several levels, stored, a directory,
    and a bzip2 member, which the reference (and this tool) report as unsupported (exit 3) unless skipped."""
    import bz2  # noqa: F401  (zipfile needs it for ZIP_BZIP2)
    rnd = os.urandom(70000)
    text = b"".join(b"line %d of some text\n" % i for i in range(20000))
    p = tmp_path / "foreign.zip"
    with zipfile.ZipFile(str(p), "w") as z:
        z.writestr(zipfile.ZipInfo("odd/"), b"")    # directory without the directory attribute: a file member to zipc
        z.mkdir("dir")
        z.writestr("dir/text1", text, compress_type=zipfile.ZIP_DEFLATED, compresslevel=1)
        z.writestr("dir/text9", text, compress_type=zipfile.ZIP_DEFLATED, compresslevel=9)
        z.writestr("rnd.stored", rnd, compress_type=zipfile.ZIP_STORED)
        z.writestr("rnd.deflated", rnd, compress_type=zipfile.ZIP_DEFLATED)
        z.writestr("empty", b"", compress_type=zipfile.ZIP_DEFLATED)
    assert zipc_tool.main(["unzip", "-t", str(p)]) == zipc_tool.EXIT_OK
    assert zipc_tool.main(["recode", "--deflate", "--level", "best", "-t", str(p)]) == zipc_tool.EXIT_OK
    q = tmp_path / "with_bz2.zip"
    with zipfile.ZipFile(str(q), "w") as z:
        z.writestr("ok", text, compress_type=zipfile.ZIP_DEFLATED)
        z.writestr("nope.bz2", text, compress_type=zipfile.ZIP_BZIP2)
    assert zipc_tool.main(["unzip", "-t", str(q)]) == zipc_tool.EXIT_UNSUPPORTED
    assert zipc_tool.main(["unzip", "-t", "-u", str(q)]) == zipc_tool.EXIT_OK
    out = tmp_path / "x"
    assert zipc_tool.main(["unzip", "-d", str(out), str(p)]) == 0
    assert (out / "dir" / "text9").read_bytes() == text and (out / "rnd.stored").read_bytes() == rnd
