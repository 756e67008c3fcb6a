"""Scratch: main bench step (8 x [encode, loss, decode] at 1024 persons) eager vs one CUDA graph."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from simple_pose_b200.pipeline import HeatmapHotPath
dev = torch.device("cuda:0")
P, B, H, W = 8192, 1024, 64, 48
nb = P // B
sets = make_inputs(P, B, H, W, dev, seed=0)
paths = [HeatmapHotPath(B, 17, H, W, device=dev) for _ in range(nb)]
def step():
    for i in range(nb):
        paths[i].step(*sets[i])
def step_fused():
    for i in range(nb):
        paths[i].train_fused(sets[i][0], sets[i][1])
        paths[i].decode(sets[i][1], sets[i][2])
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
for name, fn in (("3-kernel", step), ("fused+decode", step_fused)):
    eager = timeit(fn)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    graph = timeit(g.replay)
    os.environ["SP_NO_PDL"] = "1"
    __import__("simple_pose_b200")._abi.reload_tuning()
    nopdl = timeit(fn)
    del os.environ["SP_NO_PDL"]
    __import__("simple_pose_b200")._abi.reload_tuning()
    print("%-14s eager %.1f us  graph %.1f us  eager-no-PDL %.1f us  -> %.2f M persons/s (graph)" % (name, eager * 1e3, graph * 1e3, nopdl * 1e3, P / graph / 1e3))
