"""Per-warp timeline of the period-tiled fused kernel (scratch; needs the -DSP_TRAIN_TRACE build)."""
import ctypes, os, sys, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "scratch", "libsp_trace.so")
if not os.path.isfile(so):
    srcs = sorted(glob.glob(os.path.join(ROOT, "simple_pose_b200", "csrc", "*.cu")))
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                           "-shared", "--expt-relaxed-constexpr", "-DSP_TRAIN_TRACE", "-o", so] + srcs)
os.environ["SIMPLE_POSE_B200_LIB"] = so
import numpy as np, torch
from simple_pose_b200 import _abi, synth
from simple_pose_b200.pipeline import HeatmapHotPath
lib = _abi.lib()
lib.sp_debug_set_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
for (H, W, B) in ((64, 48, 1024), (96, 72, 512)):
    nb = 6
    hp = [HeatmapHotPath(B, 17, H, W, device=dev) for _ in range(nb)]
    jo = [synth.joints(B, height=H, width=W, seed=i, device=dev) for i in range(nb)]
    pr = [synth.heatmaps(B, height=H, width=W, seed=i, device=dev) for i in range(nb)]
    trace = torch.zeros(148 * 16 * 8, dtype=torch.int64, device=dev)
    for rep in range(3):
        for i in range(nb):
            if rep == 2 and i == nb - 1:
                lib.sp_debug_set_trace(trace.data_ptr())
            hp[i].train_fused(jo[i], pr[i])
    torch.cuda.synchronize()
    lib.sp_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(148, 16, 8).astype(np.float64)
    t0 = t[:, :, 1].min()                       # first warp past griddepcontrol.wait
    rel = lambda a: (a - t0) / 1e3
    print("=== %dx%d B=%d (us relative to the first warp released by griddepcontrol.wait)" % (H, W, B))
    print("entry (before wait): min %.1f  max %.1f" % (rel(t[:, :, 0]).min(), rel(t[:, :, 0]).max()))
    print("released:            min %.1f  median %.1f  max %.1f" % (rel(t[:, :, 1]).min(), np.median(rel(t[:, :, 1])), rel(t[:, :, 1]).max()))
    print("first chunk landed:  min %.1f  median %.1f  max %.1f" % (rel(t[:, :, 2]).min(), np.median(rel(t[:, :, 2])), rel(t[:, :, 2]).max()))
    last = rel(t[:, :, 3])
    print("warp done:           min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (last.min(), np.percentile(last, 10), np.median(last), np.percentile(last, 90), last.max()))
    cta_last = last.max(axis=1); cta_first = last.min(axis=1)
    print("per CTA last warp:   min %.1f  median %.1f  max %.1f ; spread first->last warp within CTA: median %.1f max %.1f" %
          (cta_last.min(), np.median(cta_last), cta_last.max(), np.median(cta_last - cta_first), (cta_last - cta_first).max()))
    print("after loss reduce:   min %.1f  median %.1f  max %.1f" % (rel(t[:, :, 4]).min(), np.median(rel(t[:, :, 4])), rel(t[:, :, 4]).max()))
    maps = t[:, :, 5]
    print("maps per warp: min %d max %d ; per CTA total min %d max %d" % (maps.min(), maps.max(), maps.sum(1).min(), maps.sum(1).max()))
    # utilisation profile: warps still working as a function of time
    for q in (0.5, 0.8, 0.9, 0.95, 1.0):
        tt = last.max() * q
        print("   at %.1f us: %.0f %% of warps still working" % (tt, 100.0 * (last > tt).mean()))

    smid = t[:, 0, 7].astype(int)
    order = np.argsort(cta_last)
    print("fastest CTAs (block, smid, last, drawn-ish maps):", [(int(b), int(smid[b]), round(float(cta_last[b]), 1)) for b in order[:10]])
    print("slowest CTAs:", [(int(b), int(smid[b]), round(float(cta_last[b]), 1)) for b in order[-10:]])
    vis = (jo[nb - 1][..., 2] > 0.5).reshape(-1).float().cpu().numpy()
    nm = vis.shape[0]
    drawn = np.array([vis[int(b * nm / 148):int((b + 1) * nm / 148)].sum() for b in range(148)])
    print("corr(CTA finish, visible maps in its range) = %.3f ; corr(CTA finish, blockIdx) = %.3f ; corr(finish, smid) = %.3f" %
          (np.corrcoef(cta_last, drawn)[0, 1], np.corrcoef(cta_last, np.arange(148))[0, 1], np.corrcoef(cta_last, smid)[0, 1]))
    by_sm = sorted((int(smid[b]), round(float(cta_last[b]), 1)) for b in range(148))
    print("finish by smid:", by_sm)
