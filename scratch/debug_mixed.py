import sys, torch, numpy as np
sys.path.insert(0, '.')
from oracle import heatmap_oracle as O
from simple_pose_b200.metrics.pose_metrics import GaussTaylorKeyPointDecoder
g = torch.Generator().manual_seed(3)
hm = 0.02 * torch.randn(4, 17, 64, 48, generator=g)
yy, xx = torch.meshgrid(torch.arange(64.), torch.arange(48.), indexing="ij")
for k in range(17):
    hm[0, k] = torch.exp(-((xx - 10 - k) ** 2 + (yy - 20 - k) ** 2) / 1.0) - 0.03 * (k + 1) / 17
ref_hsp, ref_max = O.gauss_taylor_decode(hm, None, return_heatmap_space=True)
base = O.argmax_coords(hm)[0]
dec = GaussTaylorKeyPointDecoder()
hsp, mx, idx = dec.decode_with_index(hm.cuda())
hsp = hsp.cpu()
d = (hsp - ref_hsp).abs().max(-1)[0]
for b in range(4):
    for k in range(17):
        if d[b, k] > 1e-4 or torch.isnan(d[b,k]):
            print(b, k, 'base', base[b, k].tolist(), 'ref', ref_hsp[b, k].tolist(), 'gpu', hsp[b, k].tolist())
