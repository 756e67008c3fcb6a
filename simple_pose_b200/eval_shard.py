"""Multi-GPU evaluation of the decode path: persons sharded across ranks, one collective.

The reference evaluates on a single device (DDP ``val()`` runs on rank 0 only,
``processors/ddp_pose_resnet_solver.py:155-156``; the DP solvers gather all heatmaps to one GPU,
``dp_pose_hrnet_solver.py:83-84,160``). Persons are independent for decode and images are
independent for OKS-NMS, so here every rank owns a contiguous range of *images* (balanced by
person count), decodes and NMS-filters its own persons with the CUDA kernels, and the only
exchange is one ``all_gather_into_tensor`` of the packed per-person results
(K*3 + 2 float32 = 212 B per person at K = 17) over NCCL/NVLink.

One process per GPU (``torchrun``); the sharding / padding / ordering logic (``shard_images``,
``person_range``, ``gather_rows``) is device-agnostic and runs under ``gloo`` on CPU tensors in the
host-side tests -- the kernels themselves (decode, NMS, ``pack_results``) need CUDA.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_images(seg_offsets, world_size):
    """Cut I images into ``world_size`` contiguous ranges with near-equal person counts.

    ``seg_offsets`` [I+1] (persons are grouped by image, ``eval.py:161-162``). Returns an int64
    array ``cuts`` [world_size+1] of image indices: rank r owns images cuts[r]..cuts[r+1]-1 and
    persons seg[cuts[r]]..seg[cuts[r+1]]-1. Never splits an image (NMS stays rank-local)."""
    seg = np.asarray(seg_offsets, dtype=np.int64)
    images = seg.shape[0] - 1
    total = int(seg[-1])
    cuts = np.zeros(world_size + 1, dtype=np.int64)
    for r in range(1, world_size):
        goal = total * r / world_size
        i = int(np.searchsorted(seg, goal, side="left"))
        if i > 0 and abs(seg[i - 1] - goal) <= abs(seg[min(i, images)] - goal):
            i -= 1
        cuts[r] = min(max(i, cuts[r - 1]), images)
    cuts[world_size] = images
    return cuts


def person_range(seg_offsets, cuts, rank):
    seg = np.asarray(seg_offsets, dtype=np.int64)
    return int(seg[cuts[rank]]), int(seg[cuts[rank + 1]])


def pack_results(coords, max_val, keep, scores):
    """[n,K,2], [n,K,1], [n] uint8, [n] float -> float32 [n, 3K+2] rows (x,y,conf)*K, keep, score:
    one launch of ``sp_pack_rows_f32`` (CUDA tensors only, like every kernel of the path)."""
    from . import _abi
    dev = _abi.require_cuda(coords, max_val, keep, scores)
    n, k = int(coords.shape[0]), int(coords.shape[1])
    row = torch.empty((n, 3 * k + 2), dtype=torch.float32, device=dev)
    c = _abi.dense(coords, torch.float32)
    m = _abi.dense(max_val, torch.float32)
    kp = _abi.dense(keep, torch.uint8)
    sc = _abi.dense(scores, torch.float64)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_pack_rows_f32(c.data_ptr(), m.data_ptr(), kp.data_ptr(), sc.data_ptr(), row.data_ptr(),
                                               n, k, _abi.stream_ptr(dev)))
    return row


def gather_rows(local_rows, counts, group=None):
    """All-gather ragged per-rank row blocks into global order.

    ``counts`` [world] = rows owned by each rank (known on every rank from the shard plan, so no
    size exchange is needed). Ranks are padded to the largest count for the fixed-size
    collective; the padding is dropped afterwards. Returns [sum(counts), C] on every rank."""
    world = dist.get_world_size(group)
    counts = [int(c) for c in counts]
    assert len(counts) == world and local_rows.shape[0] == counts[dist.get_rank(group)]
    width = local_rows.shape[1]
    longest = max(counts) if counts else 0
    if longest == 0:
        return local_rows.new_zeros((0, width))
    padded = local_rows.new_zeros((longest, width))
    padded[:local_rows.shape[0]] = local_rows
    gathered = local_rows.new_empty((world * longest, width))
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * longest:r * longest + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


class ShardedPoseEvaluator(object):
    """decode (+flip) -> rescoring -> OKS-NMS on this rank's persons, then one all-gather.

    ``plan(seg_offsets)`` once per dataset; ``run(...)`` with the heatmaps of THIS rank's
    persons (produced by this rank's backbone replica). Every rank returns the full result
    table in global person order."""

    def __init__(self, kernel_size=11, num_joints=17, in_vis_thre=0.2, oks_thre=0.9, group=None):
        from .metrics.pose_metrics import GaussTaylorKeyPointDecoder
        self.decoder = GaussTaylorKeyPointDecoder(kernel_size, num_joints)
        self.in_vis_thre, self.oks_thre = in_vis_thre, oks_thre
        self.group = group
        self.seg = None
        self.cuts = None

    def plan(self, seg_offsets):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.seg = np.asarray(seg_offsets, dtype=np.int64)
        self.cuts = shard_images(self.seg, world)
        return self.cuts

    def my_persons(self):
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        return person_range(self.seg, self.cuts, rank)

    @torch.no_grad()
    def run(self, heat_map, trans_inv, box_scores, areas, heat_map_flip=None, joint_pairs=None, boxes=None,
            input_shape=(192, 256)):
        """``boxes`` [n,4] (x1, y1, x2, y2) may replace ``trans_inv``/``areas``: both are then derived
        on the device exactly as ``BasicTransform`` does (``naive_data.box_affines``)."""
        from .datasets.naive_data import pack_keypoints, rescore_and_nms, box_affines
        if boxes is not None:
            aff = box_affines(boxes, input_shape, (int(heat_map.shape[-1]), int(heat_map.shape[-2])))
            trans_inv, areas = aff["trans_inv"], aff["area"].double()
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        lo, hi = person_range(self.seg, self.cuts, rank)
        assert heat_map.shape[0] == hi - lo, "this rank owns persons [%d, %d)" % (lo, hi)
        if heat_map_flip is None:
            coords, conf = self.decoder(heat_map, trans_inv)
        else:
            coords, conf = self.decoder.flip_call(heat_map, heat_map_flip, trans_inv, joint_pairs)
        local_seg = (self.seg[self.cuts[rank]:self.cuts[rank + 1] + 1] - lo).astype(np.int32)
        if hi > lo:
            kps = pack_keypoints(coords, conf)
            keep, scores, _ = rescore_and_nms(kps, box_scores, areas, local_seg, self.in_vis_thre, self.oks_thre)
        else:
            keep = torch.zeros(0, dtype=torch.uint8, device=coords.device)
            scores = torch.zeros(0, dtype=torch.float64, device=coords.device)
        rows = pack_results(coords, conf, keep, scores)
        if world == 1:
            return rows
        counts = [person_range(self.seg, self.cuts, r) for r in range(world)]
        return gather_rows(rows, [b - a for a, b in counts], self.group)
