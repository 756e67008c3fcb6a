import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import torch.nn.functional as F
from simple_pose_b200 import synth
from oracle import heatmap_oracle as O
dev = "cuda:0"
for seed in (1, 50):
    hm = synth.heatmaps(1024, seed=seed, noise=0.01, device=dev)
    B, K, H, W = hm.shape
    w = torch.from_numpy(O.blur_weights(11))[None, None].to(dev)
    torch.backends.cudnn.allow_tf32 = False
    bl = F.conv2d(hm.reshape(B * K, 1, H, W), w, padding=5).reshape(B, K, H, W)
    mx, idx = hm.reshape(B, K, -1).max(-1)
    y = idx // W; x = idx % W
    inner = (x > 1) & (x < W - 2) & (y > 1) & (y < H - 2) & (mx > 0)
    offs = [(0,0),(1,0),(-1,0),(0,1),(0,-1),(2,0),(-2,0),(0,2),(0,-2),(1,1),(1,-1),(-1,1),(-1,-1)]
    mins = torch.full((B, K), 1e9, device=dev); maxs = torch.full((B, K), -1e9, device=dev)
    bi = torch.arange(B, device=dev)[:, None].expand(B, K); ki = torch.arange(K, device=dev)[None].expand(B, K)
    for dx, dy in offs:
        yy = (y + dy).clamp(0, H - 1); xx = (x + dx).clamp(0, W - 1)
        v = bl[bi, ki, yy, xx]
        mins = torch.minimum(mins, v); maxs = torch.maximum(maxs, v)
    slow = inner & (mins < 2e-10)
    print(seed, "inner", inner.float().mean().item(), "slow", int(slow.sum()), "of", B * K, "mx>0", (mx > 0).float().mean().item())
    for b, k in slow.nonzero()[:6].tolist():
        print("   ", b, k, "max", mx[b, k].item(), "xy", x[b, k].item(), y[b, k].item(), "min13", mins[b, k].item(), "max13", maxs[b, k].item(), "mapmin", hm[b,k].min().item())
