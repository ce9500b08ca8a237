import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import lz4jpeg_b200 as ljb
nblk = int(sys.argv[1]) if len(sys.argv)>1 else 96
ctx = ljb.Context(0)
h = ljb.synth.random_extract(nblk*65536, seed=42)
d_all = torch.from_numpy(h).cuda()
d_out = torch.empty(2*65536, dtype=torch.uint8, device='cuda')
d_offs = torch.empty(2, dtype=torch.int64, device='cuda'); d_res = torch.zeros(3, dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
ts=[]
for b in range(nblk):
    d_in = d_all[b*65536:(b+1)*65536]
    ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
    ljb.lz4.compress_device(d_in, 65536, d_out, d_offs, d_res, ctx)
    ts.append(ctx.last_kernel_ms())
ts=np.array(ts)
print("per-block ms: mean %.3f median %.3f max %.3f min %.3f" % (ts.mean(), np.median(ts), ts.max(), ts.min()))
order=np.argsort(-ts)[:8]
print("slowest:", [(int(i), round(float(ts[i]),3)) for i in order])
print("fastest:", [(int(i), round(float(ts[i]),3)) for i in np.argsort(ts)[:5]])
np.save('/root/repo/gpurun_out/per_block_ms.npy', ts)
