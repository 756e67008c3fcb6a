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
    # single-device result computed on this rank's GPU without any collective, through the separate kernels
    solo = ShardedPoseEvaluator()
    from simple_pose_b200.datasets.naive_data import pack_keypoints, rescore_and_nms
    from simple_pose_b200.eval_shard import pack_results, row_keep, row_scores, row_keypoints, rows_equal
    c, m = solo.decoder.flip_call(hm.to(dev), hf.to(dev), tinv.to(dev))
    keep, scores, _ = rescore_and_nms(pack_keypoints(c, m), box, area, seg.numpy())
    want = pack_results(c, m, keep, scores)
    tables = []
    used = set()
    for transport, chunks in (("nccl", 1), ("nccl", 2), ("nccl", 5), ("auto", None), ("auto", 3)):
        # 5 chunks > images of some ranks: empty chunks take part in the collectives. "auto" = the fan-out transport (the
        # NMS kernel stores the rows into every rank's symmetric buffer) where torch's symmetric memory is available
        ev = ShardedPoseEvaluator(group=None, chunks=chunks, transport=transport)
        ev.plan(seg.numpy())
        lo, hi = ev.my_persons()
        for _ in range(3):                       # the transport buffers are reused by later runs
            table = ev.run(hm[lo:hi].to(dev), tinv[lo:hi].to(dev), box[lo:hi], area[lo:hi], heat_map_flip=hf[lo:hi].to(dev))
        assert table.shape == (n, 54), table.shape
        assert torch.equal(row_keypoints(table).reshape(n, 51), want[:, :51])
        assert torch.equal(row_keep(table), keep.bool())
        assert torch.equal(row_scores(table), scores)               # float64, not narrowed
        tables.append(table)
        # transport layout without the final concatenation
        raw = ev.run(hm[lo:hi].to(dev), tinv[lo:hi].to(dev), box[lo:hi], area[lo:hi], heat_map_flip=hf[lo:hi].to(dev),
                     compact=False)
        assert raw.persons == n and rows_equal(raw.rows(), table)
        used.add(ev.transport)
    assert all(rows_equal(t, tables[0]) for t in tables)
    table = tables[0]
    # every rank holds the same table
    ref = table.clone()
    dist.broadcast(ref, src=0)
    assert rows_equal(ref, table)
    dist.barrier()
    if rank == 0:
        print("sharded eval ok: %d persons, %d images, world %d, kept %d, transports %s" %
              (n, len(seg) - 1, world, int(row_keep(table).sum()), sorted(used)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
