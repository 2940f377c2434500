"""world_size-2 test of the multi-GPU host logic on CPU (gloo): units are sharded over the ranks with no
collective on the data path; per-unit results are gathered and slice checksums merged with the library's
GF(2) combine.  The per-slice checksum is computed with zlib here (there is no GPU in this container); on
the GPU box the same flow runs through zipc_b200_crc32 in bench.py --gpus N."""
import os
import socket
import zlib

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zipc_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.rand_v1(seed, n)                          # every rank can regenerate the buffer
        lo, hi = shard.slice_bounds(n, world)[rank]
        mine = torch.tensor([zlib.crc32(data[lo:hi]), zlib.adler32(data[lo:hi]), hi - lo], dtype=torch.int64)
        got = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, mine)                             # tiny results only
        # members: LPT partition, each rank handles its members, statuses gathered
        sizes = synth.member_sizes(101, seed=5)
        parts = shard.partition_lpt(sizes, world)
        mycrc = {i: zlib.crc32(synth.text_v1(100 + i, int(sizes[i]))) for i in parts[rank][:6]}
        allcrc = [None] * world
        dist.all_gather_object(allcrc, mycrc)
        if rank == 0:
            crc = shard.combine_crc32([(int(g[0]), int(g[2])) for g in got])
            ad = shard.combine_adler32([(int(g[1]), int(g[2])) for g in got])
            merged = {}
            for d in allcrc:
                merged.update(d)
            q.put((crc, ad, zlib.crc32(data), zlib.adler32(data), parts, sizes.tolist(), merged))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_combine():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, seed, world = (3 << 20) + 12345, 11, 2
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    crc, ad, want_crc, want_ad, parts, sizes, merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert crc == want_crc and ad == want_ad
    # the partition covers every member once and is balanced to within the largest member
    flat = sorted(i for part in parts for i in part)
    assert flat == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in part) for part in parts]
    assert max(loads) - min(loads) <= max(sizes)
    for i, c in merged.items():
        assert c == zlib.crc32(synth.text_v1(100 + i, int(sizes[i])))


def test_slice_bounds_and_combine_eight_ways():
    n = (1 << 22) + 777
    data = synth.rand_v1(3, n)
    for world in (1, 2, 4, 8):
        b = shard.slice_bounds(n, world)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(lo % 512 == 0 for lo, _ in b)
        assert shard.combine_crc32([(zlib.crc32(data[lo:hi]), hi - lo) for lo, hi in b]) == zlib.crc32(data)
        assert shard.combine_adler32([(zlib.adler32(data[lo:hi]), hi - lo) for lo, hi in b]) == zlib.adler32(data)
