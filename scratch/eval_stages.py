"""Stage timing of the cfg-5 eval job on one GPU (scratch): the 3-launch chain ShardedPoseEvaluator runs,
for the whole job (1-GPU shard) and for one eighth of it (the 8-GPU shard)."""
import os, sys, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simple_pose_b200 import synth, _abi
from simple_pose_b200.datasets.naive_data import box_affines
from simple_pose_b200.eval_shard import ShardedPoseEvaluator, row_keep
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
es = synth.EvalSet()
def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a.record(); out = fn(); b.record(); t1 = time.perf_counter(); b.synchronize(); ts.append(a.elapsed_time(b)); hs.append(1e3 * (t1 - t0))
    return statistics.median(ts), statistics.median(hs), out
for frac in (1, 8):
    i1 = es.images // frac
    n = int(es.seg[i1])
    seg = es.seg[:i1 + 1]
    hm = es.heatmaps(0, n, dev)
    boxes, bs = es.boxes[:n].to(dev), es.box_scores[:n].to(dev)
    print("---- %d persons, %d images" % (n, i1))
    t, h, aff = timed(lambda: box_affines(boxes, (192, 256), (48, 64))); print("box_affines        %.3f ms (host %.3f)" % (t, h))
    for chunks in (1, 2, 4):
        ev = ShardedPoseEvaluator(chunks=chunks); ev.plan(seg)
        t, h, raw = timed(lambda: ev.run(hm, None, bs, None, boxes=boxes, compact=False))
        print("evaluator chunks=%d  %.3f ms (host %.3f)  kept %d" % (chunks, t, h, int(row_keep(raw.rows()).sum())))
    ev = ShardedPoseEvaluator(chunks=1); ev.plan(seg)
    st = ev._state(dev); buf = st["buffer"]; lib = _abi.lib(); stream = _abi.stream_ptr(dev); ws = _abi.scratch(dev, stream, 16, "decode")
    blur = ev.decoder._weights_on(dev)
    dec = lambda: _abi.check(lib.sp_decode_rows_f32(hm.data_ptr(), None, None, aff["trans_inv"].data_ptr(), blur.data_ptr(), buf.data_ptr(), 54, None, None,
                                                    n, 17, 64, 48, 11, 0, ws.data_ptr(), 16, stream))
    nms = lambda: _abi.check(lib.sp_eval_rows_nms_f32(buf.data_ptr(), 54, bs.data_ptr(), None, aff["area"].data_ptr(), st["seg"][0].data_ptr(), None, None,
                                                      n, i1, 17, st["max_seg"][0], 0.2, 0.9, stream))
    t, h, _ = timed(dec); print("decode_rows        %.3f ms (host %.3f)  %.0f GB/s" % (t, h, n * 17 * 64 * 48 * 4 / t / 1e6))
    t, h, _ = timed(nms); print("eval_rows_nms      %.3f ms (host %.3f)" % (t, h))
    del hm
    torch.cuda.empty_cache()
