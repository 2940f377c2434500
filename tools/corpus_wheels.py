"""Real-world corpus check (the reference's DEVEL.md:18-30 recipe `zipc-for-each unzip -t`): every *.whl / *.zip /
*.jar under the given directories (default /opt/wheelhouse) is tested member by member on the GPU and compared
with python's zipfile (CRC of every member).  python tools/corpus_wheels.py [DIR...]"""
import glob, os, sys, time, zipfile, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zipc_b200 import zipc

dirs = sys.argv[1:] or ["/opt/wheelhouse"]
paths = sorted(p for d in dirs for ext in ("whl", "zip", "jar") for p in glob.glob(os.path.join(d, "**", "*." + ext), recursive=True))
tot_members = tot_bytes = bad = 0
t_gpu = 0.0
blobs = {p: open(p, "rb").read() for p in paths}
for p in paths:   # warm-up pass: context creation, buffer growth, first-launch costs
    r = zipc.of_binary_string(blobs[p])
    if r.is_ok():
        z = r.get_ok()
        zipc.File.to_binary_strings([m.kind for m in z.values() if m.kind is not None and zipc.File.can_extract(m.kind)])
for p in paths:
    s = blobs[p]
    r = zipc.of_binary_string(s)
    if r.is_error():
        print("PARSE-ERROR %s: %s" % (os.path.basename(p), r.message)); bad += 1; continue
    z = r.get_ok()
    ms = [z[k] for k in sorted(z)]
    files = [m for m in ms if m.kind is not None and zipc.File.can_extract(m.kind)]
    t0 = time.perf_counter()
    res = zipc.File.to_binary_strings([m.kind for m in files])
    t_gpu += time.perf_counter() - t0
    ref = zipfile.ZipFile(p)
    infos = {i.filename.encode(): i for i in ref.infolist()}
    errs = 0
    for m, rr in zip(files, res):
        info = infos.get(bytes(m.path))
        if rr.is_error() or info is None or zlib.crc32(rr.get_ok()) != info.CRC or len(rr.get_ok()) != info.file_size:
            errs += 1
            if errs <= 3:
                print("   MEMBER-ERROR %s: %s" % (bytes(m.path)[:60], rr.message if rr.is_error() else "differs from zipfile"))
        else:
            tot_bytes += len(rr.get_ok())
    tot_members += len(files)
    bad += errs > 0
    print("%-6s %5d members  %s" % ("ok" if not errs else "FAIL", len(files), os.path.basename(p)[:70]), flush=True)
print("archives %d  bad %d  members %d  bytes %.1f MB  extract time %.2f s (%.2f GB/s incl. host copies, one call per archive)"
      % (len(paths), bad, tot_members, tot_bytes / 1e6, t_gpu, tot_bytes / max(t_gpu, 1e-9) / 1e9))
sys.exit(1 if bad else 0)
