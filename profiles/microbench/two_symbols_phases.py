"""Phase cycles of the LZ4 kernel on two-symbol random data (run with LJB_LZ4_PHASES=1)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
ctx = ljb.Context(0)
n = 64 << 20
h = np.random.default_rng(0).integers(0, 2, n, dtype=np.uint8) if len(sys.argv) < 2 else np.zeros(n, np.uint8)
d_in = torch.from_numpy(h).cuda()
nb = n // 65536
d_out = torch.empty(6 * n + 4096, dtype=torch.uint8, device='cuda')
d_offs = torch.empty(nb + 1, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
for i in range(2):
    ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
    ms = ctx.last_kernel_ms()
print(f"{ms:9.2f} ms  {n/ms/1e6:8.2f} GB/s  out={int(d_res[0].item())}")
