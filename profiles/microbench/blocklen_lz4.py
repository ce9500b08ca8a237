"""LZ4 kernel throughput against the block length (the reference's own default is 300 bytes)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
ctx = ljb.Context(0)
for bl, n in ((300, 4 << 20), (1024, 8 << 20), (4096, 32 << 20), (16384, 64 << 20), (65536, 256 << 20)):
    n = n // bl * bl
    h = ljb.synth.random_extract(n, seed=42)
    d_in = torch.from_numpy(h).cuda()
    nb = n // bl
    d_out = torch.empty(2 * n + 16 * nb + 4096, dtype=torch.uint8, device='cuda')
    d_offs = torch.empty(nb + 1, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    for i in range(2):
        ljb.lz4.compress_device(d_in, bl, d_out, d_offs, d_res, ctx)
        ms = ctx.last_kernel_ms()
    print(f"block {bl:6d}: {n >> 20:4d} MiB in {ms:9.2f} ms  {n / ms / 1e6:8.3f} GB/s  ratio {int(d_res[0].item()) / n:.3f}", flush=True)
