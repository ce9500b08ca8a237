"""Host logic of the multi-GPU path, on CPU: shard arithmetic, and a world_size-2 gloo run
in which each rank encodes its shard and ONE all-gather of byte totals places the shards in the global stream."""
import os
import socket
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sh():
    import lz4jpeg_b200 as ljb

    return ljb.sharding


def test_shard_units_partition():
    sh = _sh()
    for n in (0, 1, 7, 8, 9, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                s = sh.shard_units(n, r, world)
                got.extend(range(s.first, s.first + s.count))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sh.shard_units(4, 2, 2)


def test_lz4_and_jpeg_shards():
    sh = _sh()
    n, bl = 10 * 65536 + 123, 65536
    spans = [sh.lz4_shard_bytes(n, bl, r, 4)[:2] for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(lo % bl == 0 for lo, _ in spans)
    groups = [sh.jpeg_shard_groups(16384, 16384, r, 8) for r in range(8)]
    assert sum(g.count for g in groups) == 2048 * 2048
    assert all(g.first % 2048 == 0 for g in groups)  # whole group rows
    g = [sh.jpeg_shard_groups(1200, 630, r, 2) for r in range(2)]
    assert g[0].first == 0 and g[0].count + g[1].count == 11813 and g[1].first == g[0].count
    assert sh.exclusive_bases([5, 0, 7]) == ([0, 5, 5], 12)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, block_len, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import lz4jpeg_b200 as ljb
    from oracle.pyoracle import Oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = cases.synth_text(n, seed=77)
    lo, hi, shard = ljb.sharding.lz4_shard_bytes(n, block_len, rank, world)
    # the CPU oracle stands in for the GPU encoder here (this test is about the host-side placement)
    stream, offs, _ = Oracle().lz4_compress(data[lo:hi], block_len, 1)
    body = stream[1:]  # every shard but the first omits the frame byte (ljb_lz4_compress_dev first_block > 0)
    local_total = body.size + (1 if rank == 0 else 0)
    bases, grand = ljb.sharding.gather_totals(local_total)
    q.put((rank, shard.first, shard.count, bases[rank], grand, body.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_places_shards():
    import torch.multiprocessing as mp

    from oracle.pyoracle import Oracle

    n, bl, world = 9 * 4096 + 100, 4096, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bl, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole, offs, _ = Oracle().lz4_compress(cases.synth_text(n, seed=77), bl, 1)
    out = np.zeros(res[0][4], dtype=np.uint8)
    nblocks = (n + bl - 1) // bl
    out[0] = nblocks & 0xFF  # frame byte, LZ4.c:429
    for rank, first, count, base, grand, body in res:
        start = base + (1 if rank == 0 else 0)
        out[start:start + len(body)] = np.frombuffer(body, dtype=np.uint8)
        assert grand == whole.size
    assert np.array_equal(out, whole)
    assert res[1][3] == int(offs[res[1][1]])  # rank 1's base offset == global offset of its first block
