"""GPU: the multi-GPU decomposition with the CUDA encoders.  Two processes (one per GPU where the box has two, both on
GPU 0 otherwise) encode their shards with ljb_lz4_compress_dev / ljb_jpeg_encode_rgba_dev; ONE all-gather of byte totals (NCCL with
two GPUs, gloo with one) places them; the concatenation must equal the single-GPU stream.  Plus: offsets beyond 2^32."""
import os
import socket
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, block_len, w, h, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import lz4jpeg_b200 as ljb

    two = torch.cuda.device_count() >= world
    devi = rank if two else 0
    torch.cuda.set_device(devi)
    dev = torch.device("cuda", devi)
    dist.init_process_group("nccl" if two else "gloo", rank=rank, world_size=world)
    ctx = ljb.Context(devi)
    # ---- LZ4: this rank's blocks
    data = cases.synth_text(n, seed=77)
    lo, hi, shard = ljb.sharding.lz4_shard_bytes(n, block_len, rank, world)
    nblocks = (n + block_len - 1) // block_len
    d_in = torch.from_numpy(data[lo:hi].copy()).to(dev)
    d_out = torch.empty(2 * (hi - lo) + 4096, dtype=torch.uint8, device=dev)
    d_offs = torch.empty(shard.count + 1, dtype=torch.int64, device=dev)
    d_res = torch.zeros(3, dtype=torch.int64, device=dev)
    ljb.lz4.compress_device(d_in, block_len, d_out, d_offs, d_res, ctx, first_block=shard.first, frame_blocks=nblocks)
    torch.cuda.synchronize()
    assert int(d_res[2].item()) == 0
    total = int(d_res[0].item())
    bases, grand = ljb.sharding.gather_totals(total, device=dev if two else None)
    lz = (bases[rank], grand, d_out[:total].cpu().numpy().tobytes(), d_offs.cpu().numpy().tolist())
    # ---- JPEG: this rank's group rows of one image
    img = cases.synth_image(5, w, h)
    gs = ljb.sharding.jpeg_shard_groups(w, h, rank, world)
    enc = ljb.jpeg.process(img, first_group=gs.first, ngroups=gs.count, want_coefs=False, ctx=ctx)
    jb, jgrand = ljb.sharding.gather_totals(int(enc.stream.size), device=dev if two else None)
    q.put((rank, lz, (jb[rank], jgrand, gs.first, gs.count, enc.stream.tobytes())))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def test_two_rank_gpu_shards_concatenate_to_the_single_gpu_stream():
    import torch.multiprocessing as mp

    import lz4jpeg_b200 as ljb

    n, bl, world, w, h = 37 * 4096 + 100, 4096, 2, 256, 136
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = _free_port()
    procs = [mctx.Process(target=_worker, args=(r, world, port, n, bl, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ctx = ljb.Context(0)
    try:
        whole = ljb.lz4.lz4_encode(cases.synth_text(n, seed=77), bl, ctx=ctx)
        out = np.zeros(res[0][1][1], dtype=np.uint8)
        for rank, (base, grand, body, offs), _ in res:
            assert grand == whole.stream.size
            out[base:base + len(body)] = np.frombuffer(body, dtype=np.uint8)
        assert np.array_equal(out, whole.stream)
        first1 = ljb.sharding.lz4_shard_bytes(n, bl, 1, world)[2].first
        assert res[1][1][0] == int(whole.block_offsets[first1])  # rank 1's base == global offset of its first block
        jwhole = ljb.jpeg.process(cases.synth_image(5, w, h), want_coefs=False, ctx=ctx)
        jout = np.zeros(res[0][2][1], dtype=np.uint8)
        for rank, _, (base, grand, first, count, body) in res:
            assert grand == jwhole.stream.size
            jout[base:base + len(body)] = np.frombuffer(body, dtype=np.uint8)
        assert np.array_equal(jout, jwhole.stream)
    finally:
        ctx.close()


def test_offsets_beyond_4_gib(oracle):
    """4 GiB of random hex digits in ONE launch: such text expands (many 4-byte matches at 5 bytes of sequence overhead each), so the
    stream is longer than 2^32 bytes and block offsets are 64-bit all the way.  (Uniform random BYTES would not do: a 64 KiB block
    without a single match wraps the reference's uint16 literal counter and is written as three header bytes)
    Blocks behind the 4 GiB mark are compared with the oracle; the whole table is checked for consistency."""
    import torch

    import lz4jpeg_b200 as ljb

    free, _ = torch.cuda.mem_get_info()
    if free < 14 << 30:
        pytest.skip("needs 14 GiB of device memory")
    ctx = ljb.Context(0)
    try:
        n, bl = 4 << 30, 65536
        nb = n // bl
        d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
        step = 128 << 20
        lut = torch.tensor(list(b"0123456789abcdef"), dtype=torch.uint8, device="cuda")
        for lo in range(0, n, step):  # hex digits from a 64-bit mixing function of the position
            i = torch.arange(lo, lo + step, dtype=torch.int64, device="cuda")
            z = (i ^ (i >> 30)) * -4658895280553007687
            z = (z ^ (z >> 27)) * -7723592293110705685
            d_in[lo:lo + step] = lut[(z ^ (z >> 31)) & 15]
            del i, z
        d_out = torch.empty(n + n // 4 + 32 * nb + 4096, dtype=torch.uint8, device="cuda")
        d_offs = torch.empty(nb + 1, dtype=torch.int64, device="cuda")
        d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()  # the input was generated on torch's stream; the library works on its own
        ljb.lz4.compress_device(d_in, bl, d_out, d_offs, d_res, ctx)
        torch.cuda.synchronize()
        assert int(d_res[2].item()) == 0
        total = int(d_res[0].item())
        assert total > 1 << 32
        offs = d_offs.cpu().numpy().astype(np.int64)
        assert offs[0] == 1 and offs[-1] == total and (np.diff(offs) > 0).all()
        beyond = np.nonzero(offs[:-1] > (1 << 32))[0]
        assert beyond.size > 0
        for b in (int(beyond[0]), int(beyond[beyond.size // 2]), nb - 1):
            blk = d_in[b * bl:(b + 1) * bl].cpu().numpy()
            s, _, _ = oracle.lz4_compress(blk, bl, 1)
            got = d_out[int(offs[b]):int(offs[b + 1])].cpu().numpy()
            assert np.array_equal(got, s[1:]), b
    finally:
        ctx.close()
