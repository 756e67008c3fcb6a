"""Scratch: which kernel-to-kernel transition of the bench step costs more than the sum of its parts?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs
from simple_pose_b200.pipeline import HeatmapHotPath
dev = torch.device("cuda:0")
P, B, H, W = 8192, 1024, 64, 48
nb = P // B
sets = make_inputs(P, B, H, W, dev, seed=0)
paths = [HeatmapHotPath(B, 17, H, W, device=dev) for _ in range(nb)]
E_ = lambda i: paths[i].encode(sets[i][0])
L_ = lambda i: paths[i].loss_fwd_bwd(sets[i][1])
D_ = lambda i: paths[i].decode(sets[i][1], sets[i][2])
def timeit(seq, n=40):
    def fn():
        for i in range(nb):
            for op in seq: op(i)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3
def grouped(seq, n=40):
    def fn():
        for op in seq:
            for i in range(nb): op(i)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3
for pdl in ("0",):
    os.environ["SP_NO_PDL"] = pdl
    __import__("simple_pose_b200")._abi.reload_tuning()
    r = {"E": timeit([E_]), "L": timeit([L_]), "D": timeit([D_]), "EL": timeit([E_, L_]), "LD": timeit([L_, D_]), "ED": timeit([E_, D_]),
         "DE": timeit([D_, E_]), "ELD": timeit([E_, L_, D_]), "EDL": timeit([E_, D_, L_]), "DEL": timeit([D_, E_, L_]), "grouped E*,L*,D*": grouped([E_, L_, D_]), "grouped D*,E*,L*": grouped([D_, E_, L_]), "grouped E*,D*,L*": grouped([E_, D_, L_])}
    print("SP_NO_PDL=" + pdl, {k: round(v, 1) for k, v in r.items()})
    print("   sums: E+L %.1f  L+D %.1f  E+D %.1f  E+L+D %.1f" % (r["E"] + r["L"], r["L"] + r["D"], r["E"] + r["D"], r["E"] + r["L"] + r["D"]))
