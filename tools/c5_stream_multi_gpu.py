"""C5 (BASELINE.json configs[4]) on real GPUs: ONE deflate stream spread over N GPUs.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_stream_multi_gpu.py [--gib-per-gpu 1]

The logical input is the concatenation of the ranks' slices (text-v1, seed 50 + rank).  Every rank compresses
its slice with zipc_b200_deflate_segmented (last_piece only on the last rank), inflates its own piece again with
the index (parity of bytes and CRC-32), and rank 0 checks that the concatenated pieces are ONE valid RFC 1951
stream whose CRC-32 equals zipc_b200_crc32_combine of the per-slice CRCs (zlib is the independent reader).
No collective on the data path: only (size, crc) scalars are reduced; the pieces travel to rank 0 for the
check only.  Prints one JSON line on rank 0."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib-per-gpu", type=float, default=1.0)
    ap.add_argument("--segment", type=int, default=64 << 10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from zipc_b200 import shard, synth
    from zipc_b200 import zipc_deflate as zd

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zd.Context(local)
    L = ctx.L
    n = int(args.gib_per_gpu * (1 << 30))

    def pinned(nbytes):
        p = C.c_void_p()
        assert L.zipc_b200_host_alloc(max(nbytes, 1), C.byref(p)) == 0
        return np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p.value))

    data = pinned(n)
    data[:] = synth.text_v1(50 + rank, n)
    cbuf = pinned(n + (n >> 3) + 65536)
    obuf = pinned(n)
    nmax = -(-n // args.segment) + 1
    index = np.zeros((nmax + 1, 2), dtype=np.uint64)
    ip = index.ctypes.data_as(C.POINTER(C.c_uint64))
    clen, nseg, crc, crc2, st, olen = C.c_size_t(), C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_size_t()
    last = 1 if rank == world - 1 else 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def deflate():
        return L.zipc_b200_deflate_segmented(ctx.h, 2, data.ctypes.data, n, args.segment, last, cbuf.ctypes.data, cbuf.size,
                                             C.byref(clen), ip, nmax + 1, C.byref(nseg), C.byref(crc))

    def inflate():
        return L.zipc_b200_inflate_segmented(ctx.h, cbuf.ctypes.data, clen.value, ip, nseg.value, obuf.ctypes.data, obuf.size,
                                             C.byref(olen), C.byref(crc2), C.byref(st))

    assert deflate() == 0  # warm-up (allocations)
    barrier(); t0 = time.perf_counter(); assert deflate() == 0; barrier(); td = time.perf_counter() - t0
    assert inflate() == 0 and st.value == 0
    barrier(); t0 = time.perf_counter(); assert inflate() == 0; barrier(); ti = time.perf_counter() - t0
    assert st.value == 0 and olen.value == n and crc2.value == crc.value
    assert crc.value == zlib.crc32(data), "CRC-32 of the slice"
    assert np.array_equal(obuf[:n], data), "round trip of the slice"

    # scalars: (compressed size, crc) of every rank; wall times as max over ranks
    mine = torch.tensor([clen.value, crc.value, int(td * 1e6), int(ti * 1e6)], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    allv = [v.cpu().tolist() for v in allv]
    sizes = [v[0] for v in allv]
    # the pieces go to rank 0 only to prove that their concatenation is one stream
    piece = torch.from_numpy(cbuf[:clen.value].copy()).to(f"cuda:{local}")
    pieces = None
    if world > 1:
        pad = max(sizes)
        buf = torch.zeros(pad, dtype=torch.uint8, device=f"cuda:{local}")
        buf[:clen.value] = piece
        gathered = [torch.zeros(pad, dtype=torch.uint8, device=f"cuda:{local}") for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0)
        if rank == 0:
            pieces = [g[:sizes[r]].cpu().numpy().tobytes() for r, g in enumerate(gathered)]
    else:
        pieces = [piece.cpu().numpy().tobytes()]
    if rank == 0:
        whole_crc = shard.combine_crc32([(v[1], n) for v in allv])
        d = zlib.decompressobj(-15)
        got_crc, got_len = 0, 0
        stream = b"".join(pieces)
        for off in range(0, len(stream), 64 << 20):
            out = d.decompress(stream[off:off + (64 << 20)])
            got_crc = zlib.crc32(out, got_crc); got_len += len(out)
        out = d.flush()
        got_crc = zlib.crc32(out, got_crc); got_len += len(out)
        assert d.eof, "the concatenation must end with a final block"
        assert got_len == n * world and got_crc == whole_crc, (got_len, hex(got_crc), hex(whole_crc))
        tdm, tim = max(v[2] for v in allv) / 1e6, max(v[3] for v in allv) / 1e6
        print(json.dumps({"workload": f"C5: one deflate stream of {world} x {args.gib_per_gpu} GiB text-v1 slices, {args.segment >> 10} KiB segments, level default",
                          "n_gpus": world, "uncompressed_bytes": n * world, "compressed_bytes": sum(sizes), "ratio": round(sum(sizes) / (n * world), 4),
                          "deflate_e2e_GBps": round(n * world / tdm / 1e9, 2), "inflate_e2e_GBps": round(n * world / tim / 1e9, 2),
                          "crc32": "%08x" % whole_crc, "checked": "zlib read the concatenated pieces as one stream; CRC-32 = crc32_combine of the slices"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
