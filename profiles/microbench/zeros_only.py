import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
ctx = ljb.Context(0)
n = 16 << 20
h = np.zeros(n, np.uint8)
d_in = torch.from_numpy(h).cuda()
nb = n // 65536
d_out = torch.empty(6 * n + 4096, dtype=torch.uint8, device='cuda')
d_offs = torch.empty(nb + 1, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
print(ctx.last_kernel_ms())
