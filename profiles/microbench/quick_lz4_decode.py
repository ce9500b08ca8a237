"""Timing of the LZ4 decoder, device-resident (ljb_lz4_decompress_dev) and through the host-buffer call."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
ctx = ljb.Context(0)
for seed in range(42, 50):  # a stream with a 257..259-byte match is not decodable by anyone: take a seed without one
    h = ljb.synth.random_extract(n, seed=seed)
    f = ljb.lz4.lz4_encode(h, 65536, ctx=ctx)
    print(f"seed {seed}: {n >> 20} MiB -> {f.stream.size} B, phantom sequences {f.phantom}", flush=True)
    if f.phantom == 0 or os.environ.get('LJB_ALLOW_PHANTOM'):
        break
else:
    raise SystemExit("no phantom-free stream found")
nb = f.blocks
d_comp = torch.from_numpy(f.stream).cuda()
d_offs = torch.from_numpy(f.block_offsets.astype(np.int64)).cuda()
d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
d_len = torch.empty(nb, dtype=torch.int32, device='cuda')
d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
for i in range(3):
    ljb.lz4.decompress_device(d_comp, f.stream.size, d_offs, nb, 65536, d_out, d_len, d_res, ctx)
    ms = ctx.last_kernel_ms()
    print(f"lz4 decode (device) {n >> 20} MiB: {ms:.2f} ms  {n / ms / 1e6:.1f} GB/s of output  flags={int(d_res[2].item())} bytes={int(d_res[0].item())}", flush=True)
if f.phantom == 0:
    assert np.array_equal(d_out.cpu().numpy(), h)
if os.environ.get('LJB_ALLOW_PHANTOM'):
    raise SystemExit(0)
for i in range(2):
    t = time.time()
    out = ljb.lz4.LZ4_decode(f, ctx=ctx)
    dt = time.time() - t
    print(f"lz4 decode (host buffers, pageable) {n >> 20} MiB: {dt * 1e3:.1f} ms  {n / dt / 1e9:.2f} GB/s", flush=True)
assert np.array_equal(out, h)
