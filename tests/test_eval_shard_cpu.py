"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU eval path (image-aligned
sharding, ragged all-gather, global ordering). The kernels themselves are covered by -m gpu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simple_pose_b200 import eval_shard, synth


def test_shard_images_balanced_and_image_aligned():
    _, _, _, seg = synth.nms_groups(500, mean_group=20.0, seed=1)
    seg = seg.numpy()
    for world in (1, 2, 3, 4, 8):
        cuts = eval_shard.shard_images(seg, world)
        assert cuts[0] == 0 and cuts[-1] == 500 and np.all(np.diff(cuts) >= 0)
        counts = [eval_shard.person_range(seg, cuts, r) for r in range(world)]
        assert counts[0][0] == 0 and counts[-1][1] == seg[-1]
        for (a, b), (c, d) in zip(counts[:-1], counts[1:]):
            assert b == c                                    # contiguous, nothing dropped
        sizes = np.array([b - a for a, b in counts])
        assert sizes.max() - sizes.min() <= 2 * np.diff(seg).max()      # balanced to within an image or two


def test_shard_images_degenerate():
    assert list(eval_shard.shard_images([0, 5], 4)) == [0, 0, 0, 0, 1] or list(eval_shard.shard_images([0, 5], 4))[-1] == 1
    cuts = eval_shard.shard_images([0], 2)
    assert list(cuts) == [0, 0, 0]
    cuts = eval_shard.shard_images([0, 0, 0, 7, 7], 2)
    assert cuts[-1] == 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, seg, table, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cuts = eval_shard.shard_images(seg, world)
        lo, hi = eval_shard.person_range(seg, cuts, rank)
        k = 17
        coords = table[lo:hi, :, :2].contiguous()
        conf = table[lo:hi, :, 2:].contiguous()
        keep = (torch.arange(lo, hi) % 3 == 0).to(torch.uint8)
        scores = torch.arange(lo, hi, dtype=torch.float64) * 0.5
        # row layout of eval_shard.pack_results (a CUDA kernel): (x, y, conf) * K, keep, score
        rows = torch.cat([torch.cat([coords, conf], dim=-1).reshape(hi - lo, 3 * k), keep.float()[:, None],
                          scores.float()[:, None]], dim=1)
        counts = [eval_shard.person_range(seg, cuts, r) for r in range(world)]
        full = eval_shard.gather_rows(rows, [b - a for a, b in counts])
        out[rank] = full.clone()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ragged_all_gather_over_gloo(world):
    _, _, _, seg = synth.nms_groups(23, mean_group=6.0, seed=4)
    seg = seg.numpy()
    n = int(seg[-1])
    table = torch.randn(n, 17, 3, generator=torch.Generator().manual_seed(0))
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), seg, table, out), nprocs=world, join=True)
    want = torch.cat([table.reshape(n, 51), (torch.arange(n) % 3 == 0).float()[:, None],
                      (torch.arange(n, dtype=torch.float64) * 0.5).float()[:, None]], dim=1)
    for r in range(world):
        assert torch.equal(out[r], want)                     # every rank holds the global table, in order


def test_chunk_cuts_are_image_aligned_and_cover_each_shard():
    _, _, _, seg = synth.nms_groups(300, mean_group=12.0, seed=8)
    seg = seg.numpy().astype(np.int64)
    for world in (1, 2, 5):
        cuts = eval_shard.shard_images(seg, world)
        for chunks in (1, 2, 4, 400):
            cc = eval_shard.chunk_cuts(seg, cuts, chunks)
            assert cc.shape == (world, chunks + 1)
            for r in range(world):
                assert cc[r, 0] == cuts[r] and cc[r, -1] == cuts[r + 1] and np.all(np.diff(cc[r]) >= 0)
            persons = [[int(seg[cc[r, c + 1]] - seg[cc[r, c]]) for c in range(chunks)] for r in range(world)]
            assert sum(map(sum, persons)) == seg[-1]
            if chunks <= 4:
                for r in range(world):                     # balanced to within a couple of images
                    assert max(persons[r]) - min(persons[r]) <= 3 * np.diff(seg).max()


def _chunk_worker(rank, world, port, seg, table, chunks, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the host half of ShardedPoseEvaluator.run: plan, fill this rank's slots chunk by chunk, gather each
        # chunk (asynchronously), then read the table back in global order
        ev = eval_shard.ShardedPoseEvaluator(chunks=chunks)
        ev.plan(seg)
        width = table.shape[1]
        buf = torch.zeros((chunks, world, ev.chunk_len, width))
        a = eval_shard.person_range(ev.seg, ev.cuts, rank)[0]
        handles = []
        for c, cnt in enumerate(ev.counts[rank]):
            buf[c, rank, :cnt] = table[a:a + cnt]
            a += cnt
            handles.append(eval_shard.gather_chunk(buf[c], rank, None, async_op=True))
        for h in handles:
            h.wait()
        out[rank] = eval_shard.ShardedTable(buf, ev.counts).rows().clone()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,chunks", [(2, 1), (2, 3), (3, 2)])
def test_chunked_in_place_gather_over_gloo(world, chunks):
    _, _, _, seg = synth.nms_groups(29, mean_group=5.0, seed=14)
    seg = seg.numpy()
    n = int(seg[-1])
    table = torch.randn(n, 54, generator=torch.Generator().manual_seed(1))
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_chunk_worker, args=(world, _free_port(), seg, table, chunks, out), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(out[r], table)                    # global order, padding dropped, identical on every rank


def test_row_helpers_round_trip_the_float64_score():
    rows = torch.zeros(5, 54)
    score = torch.tensor([0.1, 1e-300, 123456.789012345, 0.0, -2.5], dtype=torch.float64)
    rows[:, -2:] = score.view(torch.float32).reshape(5, 2)        # low word first (little endian), as the kernel writes
    rows[:, -3] = torch.tensor([1., 0., 1., 0., 1.])
    assert torch.equal(eval_shard.row_scores(rows), score)
    assert eval_shard.row_keep(rows).tolist() == [True, False, True, False, True]
    assert eval_shard.row_keypoints(rows).shape == (5, 17, 3) and eval_shard.row_width(17) == 54
