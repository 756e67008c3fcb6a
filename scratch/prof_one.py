"""Launch one op a few times for ncu (scratch): python scratch/prof_one.py train_fused 64 48 1024"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_pose_b200 import synth
from simple_pose_b200.pipeline import HeatmapHotPath
op, h, w, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda:0")
hp = HeatmapHotPath(batch, 17, h, w, device=dev)
joints = synth.joints(batch, height=h, width=w, seed=1, device=dev)
pred = synth.heatmaps(batch, height=h, width=w, seed=1, device=dev)
flip = synth.heatmaps(batch, height=h, width=w, seed=2, device=dev)
tinv = synth.inverse_affines(batch, height=h, width=w, seed=1, device=dev)[0]
perm = hp.decoder._perm_on(dev, 17, None)
hp.encode(joints)
torch.cuda.synchronize()
for _ in range(3):
    if op == "encode": hp.encode(joints)
    elif op == "loss": hp.loss_fwd_bwd(pred)
    elif op == "train_fused": hp.train_fused(joints, pred)
    elif op == "decode": hp.decode(pred, tinv)
    elif op == "flip_decode": hp.decode(pred, tinv, flip, perm)
torch.cuda.synchronize()
