"""Randomised CUDA-vs-oracle check over shapes the fixed parity tests do not list (scratch; B200).

    python scratch/gpu_fuzz.py [seconds] [seed]

Does not stop at the first mismatch: prints every failing case (row, shape, seed, error) and a summary."""
import os, sys, time, warnings, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import heatmap_oracle as O
from simple_pose_b200 import synth
from simple_pose_b200.commons import transforms as T
from simple_pose_b200.datasets import naive_data as ND
from simple_pose_b200.metrics import pose_metrics as PM
from simple_pose_b200.processors import loss as L

warnings.filterwarnings("ignore")
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
DEV = torch.device("cuda:0")
counts, fails = {}, []


def ulp(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def row(name, fn, *info):
    counts[name] = counts.get(name, 0) + 1
    try:
        msg = fn()
        if msg:
            fails.append((name, info, msg)); print("FAIL", name, info, msg, flush=True)
    except Exception as e:                                     # noqa: BLE001
        fails.append((name, info, repr(e))); print("EXC ", name, info, repr(e), flush=True); traceback.print_exc()


t_end = time.time() + seconds
while time.time() < t_end:
    seed = int(rng.integers(0, 2 ** 31 - 1))
    w = int(rng.integers(5, 80)); h = int(rng.integers(5, 100)); k = int(rng.integers(1, 20)); b = int(rng.integers(1, 7))
    if rng.uniform() < 0.6:
        w = (w + 3) // 4 * 4
    sigma = float(rng.choice([1.0, 2.0, 3.0]))

    def encode_case():
        j = synth.joints(b, num_joints=k, height=h, width=w, seed=seed)
        t, wt = T.encode_heat_maps(j.to(DEV), sigma, (w, h))
        ot, ow = O.encode_batch(j.numpy(), sigma, (w, h))
        d = ulp(t.cpu().numpy(), ot)
        if d.max() > 1 or (d != 0).mean() > 1e-4 or not np.array_equal(wt.cpu().numpy(), ow):
            return "max ulp %d frac %.2g weights %s" % (d.max(), (d != 0).mean(), np.array_equal(wt.cpu().numpy(), ow))
    row("encode", encode_case, b, k, h, w, sigma, seed)

    def loss_case():
        g = torch.Generator().manual_seed(seed)
        p = torch.randn(b, k, h, w, generator=g); t = torch.rand(b, k, h, w, generator=g)
        m = torch.from_numpy(rng.choice([0.0, 1.0, 1.0, 0.5], size=(b, k)).astype(np.float32))
        lo, gr = L.mse_forward_backward(p.to(DEV), t.to(DEV), m.to(DEV))
        ol, og = O.masked_mse_loss_and_grad(p, t, m)
        if abs(lo.item() - ol.item()) > 1e-5 * abs(ol.item()) + 1e-12 or not torch.allclose(gr.cpu(), og, rtol=1e-5, atol=1e-12):
            return "loss %.8g vs %.8g, grad max err %.3g" % (lo.item(), ol.item(), (gr.cpu() - og).abs().max().item())
    row("loss", loss_case, b, k, h, w, seed)

    if h >= 8 and w >= 8:
        noise = float(rng.choice([0.0, 0.005, 0.01]))
        hm = synth.heatmaps(b, joints=k, height=h, width=w, seed=seed, noise=noise)
        tinv = synth.inverse_affines(b, height=h, width=w, seed=seed)[0]

        def decode_case():
            dec = PM.GaussTaylorKeyPointDecoder(num_joints=k)
            c, m, idx = dec.decode_with_index(hm.to(DEV))
            oc, om = O.gauss_taylor_decode(hm, None, return_heatmap_space=True)
            if not torch.equal(idx.cpu().long(), O.argmax_index(hm)) or not torch.equal(m.cpu(), om):
                return "argmax/maxval mismatch"
            e = (c.cpu() - oc).abs()
            if torch.isnan(c.cpu()).ne(torch.isnan(oc)).any() or np.nanmax(e.numpy()) > 1e-4:
                return "coord err %.3g" % np.nanmax(e.numpy())
            bc, _ = PM.BasicKeyPointDecoder()(hm.to(DEV), tinv.to(DEV))
            ob, _ = O.basic_decode(hm, tinv)
            if (bc.cpu() - ob).abs().max().item() > 1e-3 * float(tinv.abs().max()):
                return "basic decode err %.3g" % (bc.cpu() - ob).abs().max().item()
        row("decode", decode_case, b, k, h, w, noise, seed)

        def flip_case():
            hf = synth.heatmaps(b, joints=k, height=h, width=w, seed=seed + 5, noise=noise)
            pairs = [[i, i + 1] for i in range(1, k - 1, 2)]
            dec = PM.GaussTaylorKeyPointDecoder(num_joints=k)
            c, m = dec.flip_call(hm.to(DEV), hf.to(DEV), synth.identity_affines(b, device=DEV), pairs)
            oc, om = O.flip_decode(hm, hf, synth.identity_affines(b), pairs)
            if not torch.equal(m.cpu(), om) or (c.cpu() - oc).abs().max().item() > 2e-3:
                return "flip err %.3g maxval %s" % ((c.cpu() - oc).abs().max().item(), torch.equal(m.cpu(), om))
        row("flip_decode", flip_case, b, k, h, w, seed)

        def fused_case():
            j = synth.joints(b, num_joints=k, height=h, width=w, seed=seed + 9)
            pred = synth.predictions_like(torch.from_numpy(O.encode_batch(j.numpy(), 2.0, (w, h))[0]), seed=seed + 1, noise=0.1)
            out = L.encode_mse_forward_backward(j.to(DEV), pred.to(DEV), want_axes=True)
            ot, ow = O.encode_batch(j.numpy(), 2.0, (w, h))
            ol, og = O.masked_mse_loss_and_grad(pred, torch.from_numpy(ot), torch.from_numpy(ow))
            if abs(out["loss"].item() - ol.item()) > 1e-5 * abs(ol.item()) + 1e-12 or not torch.allclose(out["grad"].cpu(), og, rtol=1e-5, atol=1e-12):
                return "fused loss/grad mismatch %.8g vs %.8g" % (out["loss"].item(), ol.item())
            m = torch.from_numpy(ow)[..., None, None]
            pa, _ = O.argmax_coords(pred * m); la, _ = O.argmax_coords(torch.from_numpy(ot) * m)
            if not torch.equal(out["pred_xy"].cpu(), pa) or not torch.equal(out["label_xy"].cpu(), la):
                return "fused argmax axes mismatch"
        row("fused", fused_case, b, k, h, w, seed)

        def step_case():
            # the one-launch step kernel (K = 17 layouts only: HeatmapHotPath's decoder is built per joint count)
            from simple_pose_b200.pipeline import HeatmapHotPath
            if w % 4 != 0:
                return None
            hp = HeatmapHotPath(b, k, h, w, device=DEV)
            if not hp.one_launch_supported():
                return None
            j = synth.joints(b, num_joints=k, height=h, width=w, seed=seed + 9)
            ot, ow = O.encode_batch(j.numpy(), 2.0, (w, h))
            pred = hm + (synth.predictions_like(torch.from_numpy(ot), seed=seed + 1, noise=0.005) - torch.from_numpy(ot))   # one clear peak + noise
            hp.step_one_launch(j.to(DEV), pred.to(DEV), tinv.to(DEV), with_acc=True)
            ol, og = O.masked_mse_loss_and_grad(pred, torch.from_numpy(ot), torch.from_numpy(ow))
            if abs(hp.loss.item() - ol.item()) > 1e-5 * abs(ol.item()) + 1e-12 or not torch.allclose(hp.grad.cpu(), og, rtol=1e-5, atol=1e-12):
                return "step loss/grad mismatch %.8g vs %.8g" % (hp.loss.item(), ol.item())
            d = ulp(hp.targets.cpu().numpy(), ot)
            if d.max() > 1 or (d != 0).mean() > 1e-4 or not np.array_equal(hp.weights.cpu().numpy(), ow):
                return "step targets/weights mismatch"
            oc, om = O.gauss_taylor_decode(pred, tinv)
            mag = float(tinv[:, 0, 0].abs().max())
            # coordinates are compared where the map has a real peak: the added noise turns the generator's "dead" maps
            # (all <= 0) into faint positive noise whose Taylor step is ill-conditioned (any summation order moves it);
            # bit-equality with the stand-alone decode kernel on every map is asserted in tests/test_gpu_solver_loop.py
            clear = (om > 0.1).expand_as(oc)
            err = torch.where(clear, (hp.coords.cpu() - oc).abs(), torch.zeros_like(oc))
            if not torch.equal(hp.maxval.cpu(), om) or np.nanmax(err.numpy()) > 1e-4 * mag + 2e-3:
                return "step decode err %.3g" % np.nanmax(err.numpy())
            m = torch.from_numpy(ow)[..., None, None]
            pa, _ = O.argmax_coords(pred * m); la, _ = O.argmax_coords(torch.from_numpy(ot) * m)
            if not torch.equal(hp.pred_xy.cpu(), pa) or not torch.equal(hp.label_xy.cpu(), la):
                return "step argmax axes mismatch"
        row("step", step_case, b, k, h, w, seed)

    def nms_case():
        mean_group = float(rng.choice([1.0, 6.0, 30.0, 80.0]))
        kps, box, area, seg = synth.nms_groups(int(rng.integers(1, 6)), mean_group=mean_group, seed=seed % 100003)
        keep, scores, rank = ND.rescore_and_nms(kps, box, area, seg)
        ok, osc, _ = O.rescore_and_nms(kps.numpy(), box.numpy(), area.numpy(), seg.numpy())
        if not np.array_equal(keep.cpu().numpy().astype(bool), ok) or not np.allclose(scores.cpu().numpy(), osc, rtol=1e-15, atol=0):
            return "nms keep/scores mismatch (groups of ~%g)" % mean_group
    row("nms", nms_case, seed)

    def rows_nms_case():
        # fused rescoring + NMS on float32 result rows (the float32 OKS filter with its float64 fallback)
        from simple_pose_b200 import _abi
        mean_group = float(rng.choice([1.0, 6.0, 30.0, 70.0]))
        kps, box, area, seg = synth.nms_groups(int(rng.integers(1, 6)), mean_group=mean_group, seed=seed % 100019, jitter=float(rng.choice([0.5, 2.0, 6.0])))
        k32 = kps.float()
        n = k32.shape[0]
        rows = torch.zeros(n, 54, device=DEV)
        rows[:, :51] = k32.reshape(n, 51).to(DEV)
        thr = float(rng.choice([0.9, 0.5, 0.75]))
        mx = int(np.diff(seg.numpy()).max())
        d_box, d_area, d_seg = box.to(DEV), area.to(DEV), seg.to(DEV)          # kept alive across the asynchronous launch
        _abi.check(_abi.lib().sp_eval_rows_nms_f32(rows.data_ptr(), 54, d_box.data_ptr(), d_area.data_ptr(), None, d_seg.data_ptr(),
                                                   None, None, n, len(seg) - 1, 17, mx, 0.2, thr, _abi.stream_ptr(DEV)))
        torch.cuda.synchronize()
        ok, osc, _ = O.rescore_and_nms(k32.double().numpy(), box.numpy(), area.numpy(), seg.numpy(), 0.2, thr)
        got_keep = (rows[:, 51] > 0.5).cpu().numpy()
        got_sc = rows[:, 52:54].contiguous().view(torch.float64).reshape(-1).cpu().numpy()
        if not np.array_equal(got_keep, ok) or not np.allclose(got_sc, osc, rtol=1e-15, atol=0):
            return "rows nms keep/scores mismatch (groups of ~%g, thr %g)" % (mean_group, thr)
    row("rows_nms", rows_nms_case, seed)

    def geometry_case():
        n = int(rng.integers(1, 40))
        inp = (int(rng.integers(8, 100)) * 4, int(rng.integers(8, 100)) * 4); outp = (inp[0] // 4, inp[1] // 4)
        smp = synth.train_samples(n, seed=seed % 1000003)
        geo = T.train_geometry(smp["boxes"], smp["joints"], smp["img_w"], smp["scale_ratio"], smp["rot"], smp["flip"],
                               input_shape=inp, output_shape=outp, want_input=True)
        bad = 0
        for i in range(n):
            o = O.train_sample_geometry(smp["boxes"][i].tolist(), int(smp["img_w"][i]), smp["joints"][i].numpy(), float(smp["scale_ratio"][i]),
                                        float(smp["rot"][i]), bool(smp["flip"][i]), input_shape=inp, output_shape=outp)
            if not np.allclose(geo["trans_inv_f64"][i].cpu().numpy(), o["trans_inv"], rtol=1e-9, atol=1e-9):
                bad += 1
            elif ulp(geo["joints_hm"][i].cpu().numpy(), o["joints_hm"]).max() > 1 or ulp(geo["joints_input"][i].cpu().numpy(), o["joints_input"]).max() > 1:
                bad += 1
        if bad:
            return "%d of %d persons off" % (bad, n)
    row("train_geometry", geometry_case, seed)

torch.cuda.synchronize()
print("gpu fuzz:", ", ".join("%s %d" % kv for kv in sorted(counts.items())), "| failures:", len(fails))
for f in fails[:20]:
    print("  ", f)
