"""Multi-GPU evaluation of the decode path: persons sharded across ranks, one collective.

The reference evaluates on a single device (DDP ``val()`` runs on rank 0 only,
``processors/ddp_pose_resnet_solver.py:155-156``; the DP solvers gather all heatmaps to one GPU,
``dp_pose_hrnet_solver.py:83-84,160``). Persons are independent for decode and images are
independent for OKS-NMS, so here every rank owns a contiguous range of *images* (balanced by
person count), decodes and NMS-filters its own persons with the CUDA kernels, and the only
exchange is the all-gather of the packed per-person result rows (K*3 + 3 float32 = 216 B per person at K = 17)
over NVLink. Two transports: ``fanout`` (default where torch's symmetric memory is available) -- the gather
buffer is mapped by every rank and the NMS kernel itself stores each completed row into every rank's copy
(NVLS multicast stores, or peer by peer), followed by one cross-rank barrier; ``nccl`` -- an in-place
``all_gather_into_tensor`` per chunk of the shard on NCCL's stream while the next chunk is being decoded.

Result row (``ROW_EXTRA`` = 3 trailing floats): ``(x, y, conf) * K, keep, score_lo, score_hi`` -- the last
two are the two 32-bit halves of the float64 rescored score (``row_scores`` reassembles them), so the
score reaches the COCO result file with the reference's precision (``eval.py:168-193``).

One process per GPU (``torchrun``); the sharding / padding / ordering logic (``shard_images``,
``person_range``, ``gather_rows``) is device-agnostic and runs under ``gloo`` on CPU tensors in the
host-side tests -- the kernels themselves (decode, NMS, ``pack_results``) need CUDA.
"""
import os

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


ROW_EXTRA = 3


def row_width(num_joints):
    return 3 * int(num_joints) + ROW_EXTRA


def row_keep(rows):
    """[n, 3K+3] result rows -> bool [n]: survived the OKS-NMS."""
    return rows[:, -3] > 0.5


def row_scores(rows):
    """[n, 3K+3] result rows -> float64 [n]: the rescored person score, bit-exact (two float32 slots hold
    its low and high 32 bits)."""
    return rows[:, -2:].contiguous().view(torch.float64).reshape(-1)


def row_keypoints(rows):
    """[n, 3K+3] result rows -> float32 [n, K, 3] (x, y, conf)."""
    return rows[:, :-ROW_EXTRA].reshape(rows.shape[0], -1, 3)


def rows_equal(a, b):
    """Bit-for-bit equality of two result tables. The two score slots hold halves of a float64 and may look like
    NaNs when read as float32, so ``torch.equal`` on the float view would say "different" for identical rows."""
    return a.shape == b.shape and torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def chunk_cuts(seg_offsets, cuts, chunks):
    """Per rank, cut its image range into ``chunks`` contiguous image-aligned pieces balanced by person
    count. Returns int64 [world, chunks+1] image indices (row r starts at cuts[r], ends at cuts[r+1])."""
    seg = np.asarray(seg_offsets, dtype=np.int64)
    world = len(cuts) - 1
    out = np.zeros((world, chunks + 1), dtype=np.int64)
    for r in range(world):
        lo, hi = int(cuts[r]), int(cuts[r + 1])
        sub = seg[lo:hi + 1] - seg[lo]
        out[r] = shard_images(sub, chunks) + lo
    return out


def gather_chunk(chunk_buffer, rank, group=None, async_op=False):
    """All-gather one chunk of the transport buffer: ``chunk_buffer`` [world, chunk_len, width], of which
    this rank has filled ``chunk_buffer[rank]``. On CUDA/NCCL the collective runs IN PLACE (the send slot
    is the rank's own slice of the receive buffer: no staging copy, no zero fill, no concatenation); gloo
    does not define aliased buffers, so host tensors send a copy of the slot. Returns the work handle
    (``async_op``) or None."""
    world, chunk_len, width = chunk_buffer.shape
    slot = chunk_buffer[rank]
    if not chunk_buffer.is_cuda:
        slot = slot.clone()
    return dist.all_gather_into_tensor(chunk_buffer.view(world * chunk_len, width), slot, group=group, async_op=async_op)


def _symmetric_buffer(shape, device, group, required=False):
    """A zeroed float32 buffer of ``shape`` that every rank of ``group`` has mapped into its address space (torch's
    symmetric memory: cuMem allocations exchanged between the processes, peer mappings over NVLink and, where the
    fabric supports it, an NVLS multicast mapping). Returns a dict (buffer, handle, peer_ptrs_dev, multicast) or None
    when symmetric memory cannot be set up here (then the NCCL transport is used, unless ``required``). Collective:
    every rank must call it at the same point."""
    try:
        import torch.distributed._symmetric_memory as symm_mem
        pg = group if group is not None else dist.group.WORLD
        buf = symm_mem.empty(shape, dtype=torch.float32, device=device)
        handle = symm_mem.rendezvous(buf, pg)
        buf.zero_()
        torch.cuda.synchronize(device)
        handle.barrier(channel=0)
        multicast = int(getattr(handle, "multicast_ptr", 0) or 0)
        if os.environ.get("SP_EVAL_NO_MULTICAST"):
            multicast = 0
        return {"buffer": buf, "handle": handle, "peer_ptrs_dev": int(handle.buffer_ptrs_dev), "multicast": multicast}
    except Exception as exc:
        if required:
            raise
        _symmetric_buffer.last_error = "%s: %s" % (type(exc).__name__, exc)      # why the NCCL transport was taken instead
        return None


_symmetric_buffer.last_error = None


class ShardedTable(object):
    """The all-gathered result rows in their transport layout: ``buffer`` [chunks, world, chunk_len, width]
    (rank r's c-th chunk of persons sits in ``buffer[c, r, :counts[r][c]]``; the rest of a slot is zero
    padding). ``rows()`` returns the dense [N, width] table in global person order -- a view when there is
    one rank and one chunk, otherwise one ``torch.cat`` of the valid slices."""

    def __init__(self, buffer, counts):
        self.buffer = buffer
        self.counts = [[int(c) for c in row] for row in counts]

    @property
    def persons(self):
        return sum(sum(row) for row in self.counts)

    def rows(self):
        parts = [self.buffer[c, r, :n] for r, row in enumerate(self.counts) for c, n in enumerate(row) if n > 0]
        if not parts:
            return self.buffer.new_zeros((0, self.buffer.shape[-1]))
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


class ShardedPoseEvaluator(object):
    """decode (+flip) -> rescoring + OKS-NMS on this rank's persons, chunk by chunk, each chunk all-gathered
    (in place, asynchronously) as soon as its rows are complete.

    ``plan(seg_offsets)`` once per dataset; ``run(...)`` with the heatmaps of THIS rank's persons (produced
    by this rank's backbone replica). Every rank returns the full result table in global person order.
    ``run(..., compact=False)`` returns the transport buffer without any copy (see ``ShardedTable``).
    Three launches per chunk (``sp_decode_rows_f32``, ``sp_eval_rows_nms_f32``, the NCCL all-gather) plus
    one ``sp_box_affine_f64`` per run when boxes are given; all buffers are allocated at the first run."""

    def __init__(self, kernel_size=11, num_joints=17, in_vis_thre=0.2, oks_thre=0.9, group=None, chunks=None,
                 transport="auto"):
        from .metrics.pose_metrics import GaussTaylorKeyPointDecoder
        self.decoder = GaussTaylorKeyPointDecoder(kernel_size, num_joints)
        self.num_joints = int(num_joints)
        self.in_vis_thre, self.oks_thre = float(in_vis_thre), float(oks_thre)
        self.group = group
        self.chunks = chunks
        if transport not in ("auto", "fanout", "nccl"):
            raise ValueError("transport must be 'auto', 'fanout' or 'nccl'")
        self._requested = transport
        self.transport = transport       # after the first run: the transport in use ('fanout', 'fanout-unicast', 'nccl', 'local')
        self.seg = self.cuts = self.ccuts = None
        self._dev_state = {}
        self._cur = None

    def _world_rank(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def plan(self, seg_offsets, chunks=None):
        world, _ = self._world_rank()
        self.seg = np.asarray(seg_offsets, dtype=np.int64)
        self.cuts = shard_images(self.seg, world)
        nchunks = chunks if chunks is not None else self.chunks
        if nchunks is None:
            # fan-out transport: the NMS kernel itself stores the rows into every rank's table, nothing to pipeline.
            # NCCL transport: 2 chunks, so that one chunk's all-gather hides behind the next chunk's decode
            nchunks = 1 if (world == 1 or self._requested != "nccl") else 2
        nchunks = max(1, int(nchunks))
        self.ccuts = chunk_cuts(self.seg, self.cuts, nchunks)
        self.counts = [[int(self.seg[self.ccuts[r, c + 1]] - self.seg[self.ccuts[r, c]]) for c in range(nchunks)]
                       for r in range(world)]
        self.chunk_len = max(1, max(max(row) for row in self.counts))
        self._dev_state = {}
        return self.cuts

    def my_persons(self):
        _, rank = self._world_rank()
        return person_range(self.seg, self.cuts, rank)

    def _state(self, dev):
        st = self._dev_state.get(dev)
        if st is None:
            world, rank = self._world_rank()
            nchunks = self.ccuts.shape[1] - 1
            width = row_width(self.num_joints)
            st = {"seg": [], "max_seg": [], "images": [], "symm": None, "runs": 0}
            shape = (nchunks, world, self.chunk_len, width)
            if world > 1 and self._requested in ("auto", "fanout"):
                # TWO symmetric buffers, used alternately: a fast rank's stores of run k+1 must not land in the table a
                # slow rank is still reading from run k. With two buffers a buffer is rewritten only after the barrier
                # of the run in between, which every rank reaches (in stream order) after its reads of the older table.
                first = _symmetric_buffer(shape, dev, self.group, required=self._requested == "fanout")
                if first is not None:
                    st["symm"] = [first, _symmetric_buffer(shape, dev, self.group, required=True)]
            if st["symm"] is not None:
                self.transport = "fanout" if st["symm"][0]["multicast"] else "fanout-unicast"
            else:
                st["buffer"] = torch.zeros(shape, dtype=torch.float32, device=dev)
                self.transport = "nccl" if world > 1 else "local"
            for c in range(nchunks):
                i0, i1 = int(self.ccuts[rank, c]), int(self.ccuts[rank, c + 1])
                local = (self.seg[i0:i1 + 1] - self.seg[i0]).astype(np.int32)
                st["seg"].append(torch.from_numpy(np.ascontiguousarray(local)).to(dev))
                st["max_seg"].append(int(np.diff(local).max()) if i1 > i0 else 0)
                st["images"].append(i1 - i0)
            self._dev_state[dev] = st
        return st

    # ---- incremental form: what an eval loop does (reference eval.py:133-149: the backbone emits a batch of heatmaps,
    # the decoder turns it into keypoints, nothing else is kept). begin() -> add_batch() per batch -> finish().
    def begin(self, device):
        """Start a run on ``device``: picks the transport buffer the rows of this run are written into."""
        if self.ccuts.shape[1] - 1 != 1:
            raise ValueError("the incremental form needs chunks == 1 (the NCCL transport's chunk pipeline is run()'s)")
        dev = torch.device(device)
        st = self._state(dev)
        symm = None
        if st["symm"] is not None:
            symm = st["symm"][st["runs"] & 1]
            st["runs"] += 1
        self._cur = {"dev": dev, "st": st, "symm": symm, "buf": symm["buffer"] if symm is not None else st["buffer"],
                     "area32": False}

    @torch.no_grad()
    def add_batch(self, start, heat_map, trans_inv=None, boxes=None, heat_map_flip=None, joint_pairs=None,
                  input_shape=(192, 256)):
        """Decode persons [start, start + b) of THIS rank's shard (0-based within the shard) straight into their result
        rows: one launch of ``sp_decode_rows_f32`` (preceded by ``sp_box_affine_f64`` when detection ``boxes`` [b,4]
        float64 are given instead of ``trans_inv`` [b,2,3]; the box areas are then kept for the NMS)."""
        from . import _abi
        cur = self._cur
        dev, st, buf = cur["dev"], cur["st"], cur["buf"]
        world, rank = self._world_rank()
        n = self.counts[rank][0]
        b = int(heat_map.shape[0])
        if heat_map.dim() != 4 or heat_map.shape[1] != self.num_joints or start < 0 or start + b > n:
            raise ValueError("heat_map must be [b, %d, H, W] with persons [start, start + b) inside this rank's %d" % (self.num_joints, n))
        if b == 0:
            return
        _abi.require_cuda(heat_map, heat_map_flip, trans_inv)
        k, h, w = self.num_joints, int(heat_map.shape[2]), int(heat_map.shape[3])
        hm = heat_map if (heat_map.dtype == torch.float32 and heat_map.is_contiguous()) else _abi.dense(heat_map, torch.float32)
        hf = perm = None
        if heat_map_flip is not None:
            if tuple(heat_map_flip.shape) != tuple(heat_map.shape):
                raise ValueError("heat_map_flip must have the shape of heat_map")
            hf = _abi.dense(heat_map_flip, torch.float32)
            perm = self.decoder._perm_on(dev, k, joint_pairs)
        width = buf.shape[-1]
        lib = _abi.lib()
        stream = _abi.stream_ptr(dev)
        with torch.cuda.device(dev):
            if boxes is not None:
                bx = boxes if (isinstance(boxes, torch.Tensor) and boxes.is_cuda and boxes.dtype == torch.float64
                               and boxes.is_contiguous()) else _abi.to_device(boxes, torch.float64, dev)
                if tuple(bx.shape) != (b, 4):
                    raise ValueError("boxes must be [%d, 4]" % b)
                if "tinv" not in st:
                    st["tinv"] = torch.empty((max(n, 1), 2, 3), dtype=torch.float32, device=dev)
                    st["area"] = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)
                tinv_ptr = st["tinv"].data_ptr() + 24 * start
                _abi.check(lib.sp_box_affine_f64(bx.data_ptr(), _abi.SP_BOX_XYXY, None, None, st["area"].data_ptr() + 4 * start,
                                                 tinv_ptr, None, None, b, float(input_shape[0]) / float(input_shape[1]), w, h, 1.25,
                                                 stream))
                cur["area32"] = True
                keep_alive = bx
            else:
                ti = _abi.dense(_abi.to_device(trans_inv, torch.float32, dev), torch.float32)
                if tuple(ti.shape) != (b, 2, 3):
                    raise ValueError("trans_inv must be [%d, 2, 3]" % b)
                tinv_ptr = ti.data_ptr()
                keep_alive = ti
            blur = self.decoder._weights_on(dev)
            ws = _abi.scratch(dev, stream, 16, "decode")
            slot = buf[0, rank]
            _abi.check_ws(lib.sp_decode_rows_f32(
                hm.data_ptr(), _abi.ptr(hf), _abi.ptr(perm), tinv_ptr, blur.data_ptr(), slot.data_ptr() + 4 * width * start, width,
                None, None, b, k, h, w, int(self.decoder.kernel_size), _abi.SP_DECODE_GAUSS_TAYLOR, ws.data_ptr(), ws.numel() * 8,
                stream), dev, stream)
        cur["keep_alive"] = (keep_alive, hm, hf)       # until the next call: the launches above are asynchronous

    @torch.no_grad()
    def finish(self, box_scores, areas=None, compact=True):
        """Rescoring + OKS-NMS of this rank's images on the rows written by ``add_batch`` and the exchange with the other
        ranks (fan-out from the kernel + barrier, or the in-place NCCL all-gather). ``areas`` [n] are needed unless every
        batch came with ``boxes``. Returns the table as ``run`` does."""
        from . import _abi
        cur = self._cur
        dev, st, buf, symm = cur["dev"], cur["st"], cur["buf"], cur["symm"]
        world, rank = self._world_rank()
        n = self.counts[rank][0]
        k, width = self.num_joints, buf.shape[-1]
        lib = _abi.lib()
        stream = _abi.stream_ptr(dev)
        handle = None
        with torch.cuda.device(dev):
            if n > 0:
                bs = _abi.to_device(box_scores, torch.float64, dev).reshape(-1)
                area32 = st["area"] if (cur["area32"] and areas is None) else None
                area64 = None if area32 is not None else _abi.to_device(areas, torch.float64, dev).reshape(-1)
                if bs.shape[0] != n or (area64 is not None and area64.shape[0] != n):
                    raise ValueError("box_scores / areas must describe this rank's %d persons" % n)
                slot = buf[0, rank]
                args = (slot.data_ptr(), width, bs.data_ptr(), _abi.ptr(area64), _abi.ptr(area32), st["seg"][0].data_ptr(), None, None,
                        n, st["images"][0], k, st["max_seg"][0], self.in_vis_thre, self.oks_thre)
                if symm is None:
                    _abi.check(lib.sp_eval_rows_nms_f32(*args, stream))
                else:
                    # the kernel stores every completed row into the same slot of every rank's buffer (NVLS multicast
                    # when the fabric offers it, else peer by peer): the all-gather is its epilogue
                    _abi.check(lib.sp_eval_rows_nms_fanout_f32(*args, symm["multicast"] or None, symm["peer_ptrs_dev"], world, rank,
                                                               rank * self.chunk_len * width, stream))
            if world > 1:
                if symm is not None:
                    symm["handle"].barrier(channel=0)        # every rank's rows have landed in every rank's buffer
                else:
                    handle = gather_chunk(buf[0], rank, self.group, async_op=True)
        if handle is not None:
            handle.wait()
        table = ShardedTable(buf, self.counts)
        if not compact:
            return table
        rows = table.rows()
        return rows.clone() if rows.data_ptr() == buf.data_ptr() and rows.numel() else rows

    @torch.no_grad()
    def run(self, heat_map, trans_inv, box_scores, areas, heat_map_flip=None, joint_pairs=None, boxes=None,
            input_shape=(192, 256), compact=True):
        """``boxes`` [n,4] float64 (x1, y1, x2, y2) may replace ``trans_inv``/``areas``: both are then derived
        on the device exactly as ``BasicTransform`` does (``sp_box_affine_f64``). Returns the dense
        [N, 3K+3] float32 table (``compact=True``) or the ``ShardedTable`` it would be built from.

        The host side is kept short on purpose -- at 8 GPUs a rank's whole shard decodes in ~0.4 ms, so every
        10 us of Python between the first launch and the decode launch is 2 % of the job: inputs that already
        are dense device tensors of the right dtype are used as they are, outputs are preallocated, and the
        arguments of the second kernel are prepared while the first one runs."""
        from . import _abi
        world, rank = self._world_rank()
        lo, hi = person_range(self.seg, self.cuts, rank)
        n = hi - lo
        if heat_map.dim() != 4 or heat_map.shape[0] != n or heat_map.shape[1] != self.num_joints:
            raise ValueError("this rank owns persons [%d, %d): heat_map must be [%d, %d, H, W]" % (lo, hi, n, self.num_joints))
        dev = _abi.require_cuda(heat_map, heat_map_flip, trans_inv)
        k, h, w = self.num_joints, int(heat_map.shape[2]), int(heat_map.shape[3])
        hm = heat_map if (heat_map.dtype == torch.float32 and heat_map.is_contiguous()) else _abi.dense(heat_map, torch.float32)
        hf = perm = None
        if heat_map_flip is not None:
            if tuple(heat_map_flip.shape) != tuple(heat_map.shape):
                raise ValueError("heat_map_flip must have the shape of heat_map")
            hf = _abi.dense(heat_map_flip, torch.float32)
            perm = self.decoder._perm_on(dev, k, joint_pairs)
        st = self._state(dev)
        symm = None
        if st["symm"] is not None:
            symm = st["symm"][st["runs"] & 1]
            st["runs"] += 1
        buf = symm["buffer"] if symm is not None else st["buffer"]
        width = buf.shape[-1]
        lib = _abi.lib()
        stream = _abi.stream_ptr(dev)
        area32 = area64 = None
        with torch.cuda.device(dev):
            if boxes is not None:
                bx = boxes if (isinstance(boxes, torch.Tensor) and boxes.is_cuda and boxes.dtype == torch.float64
                               and boxes.is_contiguous()) else _abi.to_device(boxes, torch.float64, dev)
                if tuple(bx.shape) != (n, 4):
                    raise ValueError("boxes must be [%d, 4]" % n)
                if "tinv" not in st:
                    st["tinv"] = torch.empty((max(n, 1), 2, 3), dtype=torch.float32, device=dev)
                    st["area"] = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)
                ti, area32 = st["tinv"], st["area"]
                _abi.check(lib.sp_box_affine_f64(bx.data_ptr(), _abi.SP_BOX_XYXY, None, None, area32.data_ptr(), ti.data_ptr(),
                                                 None, None, n, float(input_shape[0]) / float(input_shape[1]), w, h, 1.25, stream))
            else:
                ti = _abi.dense(_abi.to_device(trans_inv, torch.float32, dev), torch.float32)
                if tuple(ti.shape) != (n, 2, 3):
                    raise ValueError("trans_inv must be [%d, 2, 3]" % n)
            blur = self.decoder._weights_on(dev)
            ws = _abi.scratch(dev, stream, 16, "decode")
            ksize = int(self.decoder.kernel_size)
            map_elems = k * h * w
            bs = None
            handles = []
            a = 0
            for c, cnt in enumerate(self.counts[rank]):
                slot = buf[c, rank]
                if cnt > 0:
                    _abi.check_ws(lib.sp_decode_rows_f32(
                        hm.data_ptr() + 4 * a * map_elems, None if hf is None else hf.data_ptr() + 4 * a * map_elems,
                        _abi.ptr(perm), ti.data_ptr() + 24 * a, blur.data_ptr(), slot.data_ptr(), width, None, None,
                        cnt, k, h, w, ksize, _abi.SP_DECODE_GAUSS_TAYLOR, ws.data_ptr(), ws.numel() * 8, stream), dev, stream)
                    if bs is None:                   # prepared while the decode kernel runs
                        bs = _abi.to_device(box_scores, torch.float64, dev).reshape(-1)
                        if area32 is None:
                            area64 = _abi.to_device(areas, torch.float64, dev).reshape(-1)
                        if bs.shape[0] != n or (area64 is not None and area64.shape[0] != n):
                            raise ValueError("box_scores / areas must describe this rank's %d persons" % n)
                    if symm is None:
                        _abi.check(lib.sp_eval_rows_nms_f32(
                            slot.data_ptr(), width, bs.data_ptr() + 8 * a,
                            None if area64 is None else area64.data_ptr() + 8 * a,
                            None if area32 is None else area32.data_ptr() + 4 * a,
                            st["seg"][c].data_ptr(), None, None, cnt, st["images"][c], k, st["max_seg"][c],
                            self.in_vis_thre, self.oks_thre, stream))
                    else:
                        # the kernel stores every completed row into the same slot of every rank's buffer (NVLS
                        # multicast when the fabric offers it, else peer by peer): the all-gather is its epilogue
                        _abi.check(lib.sp_eval_rows_nms_fanout_f32(
                            slot.data_ptr(), width, bs.data_ptr() + 8 * a,
                            None if area64 is None else area64.data_ptr() + 8 * a,
                            None if area32 is None else area32.data_ptr() + 4 * a,
                            st["seg"][c].data_ptr(), None, None, cnt, st["images"][c], k, st["max_seg"][c],
                            self.in_vis_thre, self.oks_thre, symm["multicast"] or None, symm["peer_ptrs_dev"], world, rank,
                            (c * world + rank) * self.chunk_len * width, stream))
                    a += cnt
                if world > 1 and symm is None:
                    # in place: this rank's slot already lies where the collective puts it; NCCL's stream waits for
                    # the two kernels above and runs while the next chunk is being decoded on this stream
                    handles.append(gather_chunk(buf[c], rank, self.group, async_op=True))
            if world > 1 and symm is not None:
                symm["handle"].barrier(channel=0)            # every rank's rows have landed in every rank's buffer
        for hnd in handles:
            hnd.wait()                               # stream-level: later work on this stream sees the gathered rows
        table = ShardedTable(buf, self.counts)
        if not compact:
            return table                             # the transport buffer itself: valid until the next run() (fan-out
                                                     # transport: until the run after next), for readers on this stream
        rows = table.rows()
        return rows.clone() if rows.data_ptr() == buf.data_ptr() and rows.numel() else rows
