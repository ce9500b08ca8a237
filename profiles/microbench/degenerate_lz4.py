"""Timing of the LZ4 kernel on degenerate inputs (worst cases for an exact longest-match search)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
ctx = ljb.Context(0)
n = 16 << 20
rng = np.random.default_rng(0)
cases = {
    "zeros": np.zeros(n, np.uint8),
    "period2": np.tile(np.array([65, 66], np.uint8), n // 2),
    "period1000": np.tile(rng.integers(32, 127, 1000, dtype=np.uint8), n // 1000 + 1)[:n],
    "random": rng.integers(0, 256, n, dtype=np.uint8),
    "two_symbols": rng.integers(0, 2, n, dtype=np.uint8),
    "text": ljb.synth.random_extract(n, seed=1),
}
for name, h in cases.items():
    d_in = torch.from_numpy(h).cuda()
    nb = n // 65536
    d_out = torch.empty(6 * n + 4096, dtype=torch.uint8, device='cuda')
    d_offs = torch.empty(nb + 1, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    for i in range(2):
        ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
        ms = ctx.last_kernel_ms()
    print(f"{name:12s} {ms:9.2f} ms  {n/ms/1e6:8.2f} GB/s  out={int(d_res[0].item())} phantom={int(d_res[1].item())} flags={int(d_res[2].item())}", flush=True)
