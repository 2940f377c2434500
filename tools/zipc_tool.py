#!/usr/bin/env python3
"""zipc_tool -- the reference's integration harness (test/zipc_tool.ml: `zipc crc | compress | decompress |
list | unzip | recode | zip`) over libzipc_b200.

Same commands, options that matter for testing, messages and exit codes as the reference's tool
(0 ok, 1 missing path, 2 corrupted member, 3 unsupported member, 123 other error), but batch shaped: every
command hands ALL members of an archive to the GPU in one call (Zipc.File.to_binary_strings /
deflate_of_binary_strings in zipc_b200/zipc.py) instead of looping over members.

  python tools/zipc_tool.py unzip -t ARCHIVE.zip            # test every member (inflate + CRC-32)
  python tools/zipc_tool.py recode --deflate -t ARCHIVE.zip # re-compress every member, check the result in memory
  python tools/zipc_tool.py recode --deflate --level best -o OUT.zip ARCHIVE.zip
  python tools/zipc_tool.py zip -o OUT.zip FILE...          # new archive from files
  python tools/zipc_tool.py crc [--adler-32] FILE
  python tools/zipc_tool.py compress [--zlib] [--level L] IN OUT ;  decompress [--zlib] IN OUT
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

EXIT_OK, EXIT_PATH, EXIT_CORRUPTED, EXIT_UNSUPPORTED, EXIT_SOME = 0, 1, 2, 3, 123


def _read(path: str) -> bytes:
    if path == "-":
        return sys.stdin.buffer.read()
    with open(path, "rb") as f:
        return f.read()


def _write(path: str, data: bytes) -> None:
    if path == "-":
        sys.stdout.buffer.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def _log(verbose: bool, msg: str) -> None:
    if verbose:
        print(msg, file=sys.stderr)


def _members(z: dict, wanted: list[str]):
    """Members in path order; `wanted` selects by exact path or directory prefix (select_members)."""
    ms = [z[k] for k in sorted(z)]
    if not wanted:
        return ms, []
    sel, missing = [], []
    for w in wanted:
        wb = w.encode()
        hit = [m for m in ms if bytes(m.path) == wb or bytes(m.path).startswith(wb.rstrip(b"/") + b"/")]
        if not hit:
            missing.append(w)
        sel += [m for m in hit if m not in sel]
    return sel, missing


def cmd_crc(a) -> int:
    from zipc_b200 import zipc_deflate as zd
    s = _read(a.infile)
    v = zd.Adler_32.string(s) if a.adler_32 else zd.Crc_32.string(s)
    print(zd.Crc_32.pp(v))
    return EXIT_OK


def cmd_compress(a) -> int:
    from zipc_b200 import zipc_deflate as zd
    s = _read(a.infile)
    r = zd.zlib_compress(s, level=a.level) if a.zlib else zd.deflate(s, level=a.level)
    if r.is_error():
        print("%s: %s" % (a.infile, r.message), file=sys.stderr)
        return EXIT_SOME
    cs = r.get_ok()
    cs = cs[1] if a.zlib else cs
    _write(a.outfile, bytes(cs))
    _log(a.verbose, "Compressed size: %d%%" % (100 * len(cs) // max(len(s), 1)))
    return EXIT_OK


def cmd_decompress(a) -> int:
    from zipc_b200 import zipc_deflate as zd
    cs = _read(a.infile)
    r = zd.zlib_decompress(cs) if a.zlib else zd.inflate(cs)
    if r.is_error():
        print("%s: %s" % (a.infile, r.message), file=sys.stderr)
        return EXIT_SOME
    s = r.get_ok()
    s = s[0] if a.zlib else s
    _write(a.outfile, bytes(s))
    return EXIT_OK


ZIP64 = False  # --zip64: read and write ZIP64 records (beyond the reference's tool, which refuses them like the library)


def _open_archive(path: str):
    from zipc_b200 import zipc
    s = _read(path)
    r = zipc.of_binary_string(s, zip64=ZIP64)
    if r.is_error():
        print("%s: %s" % (path, r.message), file=sys.stderr)
        return None, s
    return r.get_ok(), s


def cmd_list(a) -> int:
    from zipc_b200 import zipc
    z, _ = _open_archive(a.archive)
    if z is None:
        return EXIT_SOME
    ms, missing = _members(z, a.paths)
    for w in missing:
        print("%s: %s: No such path in archive" % (a.archive, w), file=sys.stderr)
    if missing:
        return EXIT_PATH
    for m in ms:
        p = bytes(m.path).decode("utf-8", "replace")
        if a.long and m.kind is not None:
            f = m.kind
            print("%04o %10d %10d %s %s" % (m.mode, f.decompressed_size, f.compressed_size, zipc._compression_name(f.compression), p))
        else:
            print(p)
    return EXIT_OK


def _check_members(archive: str, ms, verbose: bool, skip: bool) -> int:
    """check_archive (zipc_tool.ml:613-633), all extractable members in one GPU call."""
    from zipc_b200 import zipc
    files = [m for m in ms if m.kind is not None and zipc.File.can_extract(m.kind)]
    res = dict(zip((id(m) for m in files), zipc.File.to_binary_strings([m.kind for m in files])))
    code = EXIT_OK
    for m in ms:
        p = bytes(m.path).decode("utf-8", "replace")
        if m.kind is None:
            _log(verbose, "[----] %s" % p)
        elif id(m) in res:
            r = res[id(m)]
            if r.is_error():
                _log(verbose, "[FAIL] %s %s" % (p, r.message))
                code = EXIT_CORRUPTED
            else:
                _log(verbose, "[ OK ] %s" % p)
        else:
            enc = " encrypted" if zipc.File.is_encrypted(m.kind) else ""
            _log(verbose, "[ ?? ] %s%s %s" % (p, enc, zipc._compression_name(m.kind.compression)))
            if not skip:
                print("%s: %s: unsupported compression or encryption" % (archive, p), file=sys.stderr)
                code = max(code, EXIT_UNSUPPORTED) if code != EXIT_CORRUPTED else code
    if code != EXIT_OK:
        print("%s: Some archive members had errors" % archive, file=sys.stderr)
    else:
        _log(verbose, "\nNo errors in %s" % archive)
    return code


def cmd_unzip(a) -> int:
    from zipc_b200 import zipc
    z, _ = _open_archive(a.archive)
    if z is None:
        return EXIT_SOME
    ms, missing = _members(z, a.paths)
    for w in missing:
        print("%s: %s: No such path in archive" % (a.archive, w), file=sys.stderr)
    if missing:
        return EXIT_PATH
    if a.test:
        return _check_members(a.archive, ms, a.verbose, a.skip_unsupported)
    root = a.directory or os.path.splitext(os.path.basename(a.archive))[0]
    if os.path.exists(root) and not a.force:
        print("%s: Directory exists" % root, file=sys.stderr)
        return EXIT_SOME
    files = [m for m in ms if m.kind is not None and zipc.File.can_extract(m.kind)]
    datas = zipc.File.to_binary_strings([m.kind for m in files])
    code = EXIT_OK
    for m in ms:
        if m.kind is None:
            os.makedirs(os.path.join(root, bytes(m.path).decode()), exist_ok=True)
    for m, r in zip(files, datas):
        p = bytes(m.path).decode("utf-8", "replace")
        if p.startswith("/") or ".." in p.split("/"):
            print("%s: %s: absurd path, skipped" % (a.archive, p), file=sys.stderr)
            continue
        if r.is_error():
            print("%s: %s" % (p, r.message), file=sys.stderr)
            code = EXIT_CORRUPTED
            continue
        dst = os.path.join(root, p)
        if p.endswith("/"):  # a directory written as an empty file member (no directory attribute): some zippers do
            os.makedirs(dst, exist_ok=True)
            continue
        os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
        with open(dst, "wb") as f:
            f.write(r.get_ok())
        os.chmod(dst, m.mode & 0o777)
        _log(a.verbose, p)
    return code


def _recode(z: dict, ms, compression: str | None, level: str, verbose: bool):
    """recode_members (zipc_tool.ml:440-465): extract all, re-compress all, two GPU calls."""
    from zipc_b200 import zipc, zipc_deflate as zd
    if compression is None:
        return z, None
    files = [m for m in ms if m.kind is not None and zipc.File.can_extract(m.kind)]
    datas = zipc.File.to_binary_strings([m.kind for m in files])
    for m, r in zip(files, datas):
        if r.is_error():
            return None, "%s: %s" % (bytes(m.path).decode("utf-8", "replace"), r.message)
    payloads = [r.get_ok() for r in datas]
    if compression == "deflate":
        new = zipc.File.deflate_of_binary_strings(payloads, level)
    else:
        new = [zipc.File.stored_of_binary_string(p) for p in payloads]
    for m, d, r in zip(files, payloads, new):
        p = bytes(m.path).decode("utf-8", "replace")
        if r.is_error():
            return None, "%s: %s" % (p, r.message)
        f2 = r.get_ok()
        if f2.decompressed_crc_32 != m.kind.decompressed_crc_32:
            return None, "%s: Recoding changed the checksum from %s to %s (zipc bug)" % (
                p, zd.Crc_32.pp(m.kind.decompressed_crc_32), zd.Crc_32.pp(f2.decompressed_crc_32))
        m2 = zipc.Member.make(m.path, f2, mode=m.mode, mtime=m.mtime).get_ok()
        old = 100 * m.kind.compressed_size // max(m.kind.decompressed_size, 1)
        _log(verbose, "Recode %3d%% (was %3d%%) %s" % (100 * f2.compressed_size // max(f2.decompressed_size, 1), old, p))
        z = zipc.add(m2, z)
    return z, None


def cmd_recode(a) -> int:
    from zipc_b200 import zipc
    z, s = _open_archive(a.archive)
    if z is None:
        return EXIT_SOME
    ms, missing = _members(z, a.paths)
    for w in missing:
        print("%s: %s: No such path in archive" % (a.archive, w), file=sys.stderr)
    if missing:
        return EXIT_PATH
    comp = "deflate" if a.deflate else "stored" if a.stored else None
    z2, err = _recode(z, ms, comp, a.level, a.verbose)
    if err:
        print(err, file=sys.stderr)
        return EXIT_SOME
    r = zipc.to_binary_string(z2, zip64=ZIP64)
    if r.is_error():
        print("%s: %s" % (a.archive, r.message), file=sys.stderr)
        return EXIT_SOME
    out = r.get_ok()
    if a.test:  # recode_check_in_memory (zipc_tool.ml:498-515)
        r2 = zipc.of_binary_string(out, zip64=ZIP64)
        if r2.is_error():
            print("recode check: %s" % r2.message, file=sys.stderr)
            return EXIT_SOME
        z3 = r2.get_ok()
        code = _check_members(a.archive, [z3[k] for k in sorted(z3)], False, True)
        if code == EXIT_OK:
            _log(a.verbose, "No errors in %s recode (%d%% of old size)" % (a.archive, 100 * len(out) // max(len(s), 1)))
        return code
    _write(a.output, out)
    return EXIT_OK


def cmd_zip(a) -> int:
    from zipc_b200 import zipc
    paths, payloads, modes, mtimes = [], [], [], []
    for p in a.files:
        if os.path.isdir(p):
            for root, _, fs in os.walk(p):
                for f in sorted(fs):
                    paths.append(os.path.join(root, f))
        elif os.path.exists(p):
            paths.append(p)
        else:
            print("%s: No such file or directory" % p, file=sys.stderr)
            return EXIT_PATH
    for p in paths:
        st = os.stat(p)
        payloads.append(_read(p)); modes.append(st.st_mode & 0o777); mtimes.append(int(st.st_mtime))
    names = [os.path.relpath(p, a.strip_prefix) if a.strip_prefix else p for p in paths]
    if a.stored:
        z = zipc.empty()
        for nm, d, md, mt in zip(names, payloads, modes, mtimes):
            f = zipc.File.stored_of_binary_string(d).get_ok()
            z = zipc.add(zipc.Member.make(nm, f, mode=md, mtime=mt).get_ok(), z)
        r = zipc.to_binary_string(z, zip64=ZIP64)
    else:
        r = zipc.archive_of_binary_strings(names, payloads, a.level, modes, mtimes, zip64=ZIP64)
    if r.is_error():
        print(r.message, file=sys.stderr)
        return EXIT_SOME
    _write(a.output, r.get_ok())
    return EXIT_OK


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="zipc_tool", description=__doc__.split("\n\n")[0])
    ap.add_argument("--zip64", action="store_true", help="accept and write ZIP64 archives (the reference refuses them)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    levels = ["none", "fast", "default", "best"]
    p = sub.add_parser("crc"); p.add_argument("--adler-32", action="store_true"); p.add_argument("infile"); p.set_defaults(fn=cmd_crc)
    for name, fn in (("compress", cmd_compress), ("decompress", cmd_decompress)):
        p = sub.add_parser(name); p.add_argument("--zlib", action="store_true"); p.add_argument("-v", "--verbose", action="store_true")
        if name == "compress":
            p.add_argument("--level", choices=levels, default="default")
        p.add_argument("infile"); p.add_argument("outfile"); p.set_defaults(fn=fn)
    p = sub.add_parser("list"); p.add_argument("-l", "--long", action="store_true"); p.add_argument("archive"); p.add_argument("paths", nargs="*"); p.set_defaults(fn=cmd_list)
    p = sub.add_parser("unzip"); p.add_argument("-t", "--test", action="store_true"); p.add_argument("-v", "--verbose", action="store_true")
    p.add_argument("-u", "--skip-unsupported", action="store_true"); p.add_argument("-d", "--directory"); p.add_argument("-f", "--force", action="store_true")
    p.add_argument("archive"); p.add_argument("paths", nargs="*"); p.set_defaults(fn=cmd_unzip)
    p = sub.add_parser("recode"); g = p.add_mutually_exclusive_group(); g.add_argument("--deflate", action="store_true"); g.add_argument("--stored", action="store_true")
    p.add_argument("--level", choices=levels, default="default"); p.add_argument("-t", "--test", action="store_true"); p.add_argument("-v", "--verbose", action="store_true")
    p.add_argument("-o", "--output", default="-"); p.add_argument("archive"); p.add_argument("paths", nargs="*"); p.set_defaults(fn=cmd_recode)
    p = sub.add_parser("zip"); p.add_argument("-o", "--output", required=True); p.add_argument("--stored", action="store_true")
    p.add_argument("--level", choices=levels, default="default"); p.add_argument("--strip-prefix"); p.add_argument("files", nargs="+"); p.set_defaults(fn=cmd_zip)
    a = ap.parse_args(argv)
    global ZIP64
    ZIP64 = a.zip64
    try:
        return a.fn(a)
    except OSError as e:
        print("%s" % e, file=sys.stderr)
        return EXIT_SOME


if __name__ == "__main__":
    sys.exit(main())
