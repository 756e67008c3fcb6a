"""Does the loss kernel find the encoder's targets in L2 when it follows it directly? (scratch)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_pose_b200 import synth
from simple_pose_b200.pipeline import HeatmapHotPath
dev = torch.device("cuda:0")
H, W = 64, 48
for B in (1024, 512, 256, 128):
    nb = max(8, 8192 // B)
    hp = [HeatmapHotPath(B, 17, H, W, device=dev) for _ in range(nb)]
    jo = [synth.joints(B, seed=i, device=dev) for i in range(nb)]
    pr = [synth.heatmaps(B, seed=i, device=dev) for i in range(nb)]
    def grouped():
        for i in range(nb): hp[i].encode(jo[i])
        for i in range(nb): hp[i].loss_fwd_bwd(pr[i])
    def paired():
        for i in range(nb):
            hp[i].encode(jo[i]); hp[i].loss_fwd_bwd(pr[i])
    for name, fn in (("grouped", grouped), ("paired", paired)):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        print("B=%4d %-8s %8.1f us per %d persons -> %6.1f us per 1024 persons" % (B, name, ms * 1e3, nb * B, ms * 1e3 * 1024 / (nb * B)), flush=True)
    del hp, jo, pr
    torch.cuda.empty_cache()
