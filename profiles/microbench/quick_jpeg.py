"""Time (and, under ncu, profile) the JPEG encode kernel alone on a seed-42 noise image of DIM x DIM."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
hh = int(sys.argv[2]) if len(sys.argv) > 2 else dim
bpp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = ljb.Context(0)
img = np.ascontiguousarray(ljb.synth.random_image(dim, hh, seed=42)[:, :, :bpp])
d_in = torch.from_numpy(img).cuda()
ng = ljb.jpeg.group_count(dim, hh)
d_out = torch.empty(ng * 96 + 4096, dtype=torch.uint8, device='cuda')
d_offs = torch.empty(ng + 1, dtype=torch.int64, device='cuda')
d_bits = torch.empty(ng * 3, dtype=torch.int16, device='cuda')
d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
for i in range(3):
    ljb.jpeg.encode_device(d_in, dim, hh, d_out, d_offs, d_bits, d_res, ctx, bpp=bpp)
    ms = ctx.last_kernel_ms()
    print(f"jpeg {dim}x{hh} bpp={bpp}: {ms:.3f} ms  {dim*hh/ms/1e3:.1f} MPix/s  out={int(d_res[0].item())} flags={int(d_res[2].item())}")
