"""Per-CTA timeline of the loss kernel (scratch; -DSP_TRAIN_TRACE build)."""
import ctypes, os, sys, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "scratch", "libsp_trace.so")
srcs = sorted(glob.glob(os.path.join(ROOT, "simple_pose_b200", "csrc", "*.cu")))
if not os.path.isfile(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                           "-shared", "--expt-relaxed-constexpr", "-DSP_TRAIN_TRACE", "-o", so] + srcs)
os.environ["SIMPLE_POSE_B200_LIB"] = so
import numpy as np, torch
from simple_pose_b200 import _abi, synth
from simple_pose_b200.pipeline import HeatmapHotPath
lib = _abi.lib()
lib.sp_debug_set_loss_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
for (H, W, B) in ((64, 48, 1024), (96, 72, 512)):
    nb = 6
    hp = [HeatmapHotPath(B, 17, H, W, device=dev) for _ in range(nb)]
    jo = [synth.joints(B, height=H, width=W, seed=i, device=dev) for i in range(nb)]
    pr = [synth.heatmaps(B, height=H, width=W, seed=i, device=dev) for i in range(nb)]
    for i in range(nb):
        hp[i].encode(jo[i])
    trace = torch.zeros(148 * 32 * 4, dtype=torch.int64, device=dev)
    for rep in range(3):
        for i in range(nb):
            if rep == 2 and i == nb - 1:
                lib.sp_debug_set_loss_trace(trace.data_ptr())
            hp[i].loss_fwd_bwd(pr[i])
    torch.cuda.synchronize()
    lib.sp_debug_set_loss_trace(None)
    t = trace.cpu().numpy().reshape(148, 32, 4).astype(np.float64)
    used = t[:, :, 1] > 0
    t0 = t[:, :, 0][used].min()
    done = np.where(used, (t[:, :, 1] - t0) / 1e3, np.nan)
    cta_last = np.nanmax(done, axis=1)
    smid = t[:, 0, 2].astype(int)
    print("=== loss %dx%d B=%d: warps/CTA %d" % (H, W, B, int(used[0].sum())))
    print("warp done: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (np.nanmin(done), np.nanpercentile(done, 10), np.nanmedian(done), np.nanpercentile(done, 90), np.nanmax(done)))
    print("CTA last warp: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (cta_last.min(), np.percentile(cta_last, 10), np.median(cta_last), np.percentile(cta_last, 90), cta_last.max()))
    by = sorted((int(smid[b]), round(float(cta_last[b]), 1)) for b in range(148))
    print("by smid:", by[:48])
