"""Launches every hot-path kernel configuration a few times for ncu (see profiles/README.md).

    ncu --set full --clock-control none --import-source on -k regex:"decode_|encode_|mse_fwd|oks_" \
        -o gpurun_out/rN_prof python profiles/prof_driver.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple_pose_b200 import synth  # noqa: E402
from simple_pose_b200.pipeline import HeatmapHotPath  # noqa: E402
from simple_pose_b200.datasets.naive_data import rescore_and_nms  # noqa: E402

dev = torch.device("cuda:0")
reps = int(os.environ.get("PROF_REPS", "2"))
for (h, w, batch) in ((64, 48, 1024), (96, 72, 512)):  # summarize.py labels launches by this order
    hp = HeatmapHotPath(batch, 17, h, w, device=dev)
    joints = synth.joints(batch, height=h, width=w, seed=1, device=dev)
    pred = synth.heatmaps(batch, height=h, width=w, seed=1, device=dev)
    flip = synth.heatmaps(batch, height=h, width=w, seed=2, device=dev)
    tinv = synth.inverse_affines(batch, height=h, width=w, seed=1, device=dev)[0]
    perm = hp.decoder._perm_on(dev, 17, None)
    torch.cuda.synchronize()
    for _ in range(reps):
        hp.encode(joints)
        hp.loss_fwd_bwd(pred)
        hp.decode(pred, tinv)
        hp.decode(pred, tinv, flip, perm)
        hp.train_fused(joints, pred)
        hp.step_one_launch(joints, pred, tinv)
    torch.cuda.synchronize()
# the literal batch-128 step (BASELINE configs 1 + 2): one launch
hp = HeatmapHotPath(128, 17, 64, 48, device=dev)
joints = synth.joints(128, seed=3, device=dev)
pred = synth.heatmaps(128, seed=3, device=dev)
tinv = synth.inverse_affines(128, seed=3, device=dev)[0]
for _ in range(reps):
    hp.step_one_launch(joints, pred, tinv)
torch.cuda.synchronize()
from simple_pose_b200.commons.transforms import train_geometry  # noqa: E402
smp = {k: v.to(dev) for k, v in synth.train_samples(8192, seed=5).items()}
for _ in range(reps):
    train_geometry(smp["boxes"], smp["joints"], smp["img_w"], smp["scale_ratio"], smp["rot"], smp["flip"].to(torch.uint8))
torch.cuda.synchronize()
kps, box, area, seg = synth.nms_groups(512, mean_group=20.0, seed=3)
for _ in range(reps):
    rescore_and_nms(kps.to(dev), box.to(dev), area.to(dev), seg)
torch.cuda.synchronize()
# the eval chain of one rank of an 8-GPU run of BASELINE config 5 (an eighth of the job): box -> affine, decode into
# the result rows, fused rescoring + OKS-NMS on the rows
from simple_pose_b200.eval_shard import ShardedPoseEvaluator  # noqa: E402
es = synth.EvalSet()
i1 = es.images // 8
n = int(es.seg[i1])
ev = ShardedPoseEvaluator(chunks=1)
ev.plan(es.seg[:i1 + 1])
hm = es.heatmaps(0, n, dev)
boxes, bs = es.boxes[:n].to(dev), es.box_scores[:n].to(dev)
torch.cuda.synchronize()
for _ in range(reps):
    ev.run(hm, None, bs, None, boxes=boxes, compact=False)
torch.cuda.synchronize()
print("done")
