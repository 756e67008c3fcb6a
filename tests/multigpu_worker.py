"""torchrun worker for tests/test_multigpu_gpu.py (not collected by pytest)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from simple_pose_b200 import synth  # noqa: E402
from simple_pose_b200.eval_shard import ShardedPoseEvaluator  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    # the whole synthetic eval set, identical on every rank (seeded); each rank uses its shard
    _, box, _, seg = synth.nms_groups(37, mean_group=9.0, seed=5)
    n = int(seg[-1])
    hm, hf = synth.flip_pair(n, seed=6)
    tinv, area = synth.inverse_affines(n, seed=6)
    ev = ShardedPoseEvaluator(group=None)
    ev.plan(seg.numpy())
    lo, hi = ev.my_persons()
    table = ev.run(hm[lo:hi].to(dev), tinv[lo:hi].to(dev), box[lo:hi], area[lo:hi], heat_map_flip=hf[lo:hi].to(dev))
    assert table.shape == (n, 53), table.shape
    # single-device result computed on this rank's GPU without any collective
    solo = ShardedPoseEvaluator()
    solo.seg, solo.cuts = seg.numpy().astype(np.int64), np.array([0, len(seg) - 1], dtype=np.int64)
    from simple_pose_b200.datasets.naive_data import pack_keypoints, rescore_and_nms
    from simple_pose_b200.eval_shard import pack_results
    c, m = solo.decoder.flip_call(hm.to(dev), hf.to(dev), tinv.to(dev))
    keep, scores, _ = rescore_and_nms(pack_keypoints(c, m), box, area, seg.numpy())
    want = pack_results(c, m, keep, scores)
    assert torch.equal(table, want), (table - want).abs().max().item()
    # every rank holds the same table
    ref = table.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, table)
    dist.barrier()
    if rank == 0:
        print("sharded eval ok: %d persons, %d images, world %d, kept %d" % (n, len(seg) - 1, world, int(table[:, 51].sum())))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
