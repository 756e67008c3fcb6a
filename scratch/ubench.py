"""Micro-benchmark of single kernels over rotating buffers (scratch tool, not part of the product).

    python scratch/ubench.py --ops decode,encode --batch 1024,4096 --hw 64x48 [--env "A=1,B=2" --env ""]

Each (--env) set is applied to os.environ before the timed loop (the library reads its tuning
knobs with getenv at every call).
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple_pose_b200 import synth, _abi  # noqa: E402
from simple_pose_b200.pipeline import HeatmapHotPath, ALGO_BYTES  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="encode,loss,train_fused,decode,flip_decode")
ap.add_argument("--batch", default="1024")
ap.add_argument("--hw", default="64x48")
ap.add_argument("--total-mb", type=int, default=1700, help="pred bytes across all rotating buffer sets")
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--env", action="append", default=None)
ap.add_argument("--peak", type=float, default=6537.0)
args = ap.parse_args()
dev = torch.device("cuda:0")
envs = args.env if args.env else [""]

for hw in args.hw.split(","):
    H, W = [int(x) for x in hw.split("x")]
    for batch in [int(b) for b in args.batch.split(",")]:
        per = 17 * H * W * 4 * batch
        nb = max(2, min(64, (args.total_mb << 20) // per))
        gen = min(batch, 1024)
        sets, flips, paths = [], [], []
        for i in range(nb):
            rep = batch // gen
            j = synth.joints(gen, height=H, width=W, seed=7 + i, device=dev).repeat(rep, 1, 1)
            p = synth.heatmaps(gen, height=H, width=W, seed=7 + i, noise=0.01, device=dev).repeat(rep, 1, 1, 1)
            t = synth.inverse_affines(gen, height=H, width=W, seed=7 + i, device=dev)[0].repeat(rep, 1, 1)
            f = synth.heatmaps(gen, height=H, width=W, seed=900 + i, noise=0.01, device=dev).repeat(rep, 1, 1, 1)
            sets.append((j, p, t))
            flips.append(f)
            paths.append(HeatmapHotPath(batch, 17, H, W, device=dev))
        perm = paths[0].decoder._perm_on(dev, 17, None)
        ops = {
            "encode": lambda i: paths[i].encode(sets[i][0]),
            "loss": lambda i: paths[i].loss_fwd_bwd(sets[i][1]),
            "train_fused": lambda i: paths[i].train_fused(sets[i][0], sets[i][1]),
            "decode": lambda i: paths[i].decode(sets[i][1], sets[i][2]),
            "flip_decode": lambda i: paths[i].decode(sets[i][1], sets[i][2], flips[i], perm),
            "step": lambda i: paths[i].step_one_launch(sets[i][0], sets[i][1], sets[i][2]),
            "step_train": lambda i: paths[i].step_one_launch(sets[i][0], sets[i][1], None, want_targets=False, with_acc=True, want_decode=False),
        }
        for i in range(nb):
            paths[i].encode(sets[i][0])      # targets for the loss
        for env in envs:
            saved = {}
            for kv in [x for x in env.split(",") if x]:
                k, v = kv.split("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = v
            _abi.reload_tuning()
            for name in args.ops.split(","):
                fn = ops[name]
                for i in range(nb):
                    fn(i)
                torch.cuda.synchronize()
                times = []
                for _ in range(args.reps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(nb):
                        fn(i)
                    b.record()
                    b.synchronize()
                    times.append(a.elapsed_time(b) / nb)
                ms = statistics.median(times)
                best = min(times)
                by = ALGO_BYTES[{'step_train': 'train_fused'}.get(name, name)](17, H, W) * batch
                gbs = by / (ms * 1e-3) / 1e9
                print("%-12s %dx%d B=%-6d nb=%-2d env=[%s]  %8.2f us (best %8.2f)  %7.1f GB/s  frac %.3f" %
                      (name, H, W, batch, nb, env, ms * 1e3, best * 1e3, gbs, gbs / args.peak), flush=True)
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            _abi.reload_tuning()
        del sets, flips, paths
        torch.cuda.empty_cache()
