"""Stage timing of the cfg-5 eval job on one GPU (scratch)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simple_pose_b200 import synth
from simple_pose_b200.metrics.pose_metrics import GaussTaylorKeyPointDecoder
from simple_pose_b200.datasets.naive_data import pack_keypoints, rescore, oks_nms_batched, box_affines
from simple_pose_b200.eval_shard import pack_results
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
persons, mean_group = 104000, 20.0
images = int(persons / (1.0 + mean_group))
sizes = 1 + torch.poisson(torch.full((images,), mean_group), generator=g).long()
seg = np.zeros(images + 1, dtype=np.int64); seg[1:] = np.cumsum(sizes.numpy()); n = int(seg[-1])
hm = torch.empty((n, 17, 64, 48), dtype=torch.float32, device=dev)
for a in range(0, n, 8192):
    b = min(n, a + 8192); hm[a:b] = synth.heatmaps(b - a, seed=777 + a, device=dev)
boxes = synth.detection_boxes(n, seed=778).to(dev)
box_scores = ((torch.randperm(n, generator=g).double() + 0.5) / n).to(dev)
dec = GaussTaylorKeyPointDecoder()
seg32 = seg.astype(np.int32)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts), out
t, aff = timed(lambda: box_affines(boxes, (192, 256), (48, 64))); print("box_affines %.3f ms" % t)
areas = aff["area"].double()
t, (coords, conf) = timed(lambda: dec(hm, aff["trans_inv"])); print("decode %.3f ms" % t)
t, kps = timed(lambda: pack_keypoints(coords, conf)); print("pack_keypoints %.3f ms" % t)
t, scores = timed(lambda: rescore(kps, box_scores, 0.2)); print("rescore %.3f ms" % t)
t, (keep, rank) = timed(lambda: oks_nms_batched(kps, scores, areas, seg32, 0.9)); print("oks_nms %.3f ms (max seg %d)" % (t, int(sizes.max())))
t, rows = timed(lambda: pack_results(coords, conf, keep, scores)); print("pack_results %.3f ms" % t)
print("kept", int(keep.sum().item()), "of", n)
