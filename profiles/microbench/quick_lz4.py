import sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import lz4jpeg_b200 as ljb
n = int(sys.argv[1]) if len(sys.argv)>1 else 256<<20
ctx = ljb.Context(0)
h = ljb.synth.random_extract(n, seed=42)
d_in = torch.from_numpy(h).cuda()
nb = n//65536
d_out = torch.empty(n + n//8 + 16*nb, dtype=torch.uint8, device='cuda')
d_offs = torch.empty(nb+1, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
for i in range(3):
    ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
    ms = ctx.last_kernel_ms()
    print(f"lz4 {n>>20} MiB: {ms:.2f} ms  {n/ms/1e6:.2f} GB/s  out={int(d_res[0].item())}")
