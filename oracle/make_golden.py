"""TEST INFRASTRUCTURE ONLY -- freezes outputs of the *reference itself* as fixtures.

Run in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden

It imports the unmodified reference through ``oracle/ref_loader.py`` (two documented
shims), feeds it deterministic inputs, and writes ``tests/golden/*.npz``. The fixtures
travel to the GPU box; the reference tree does not. Inputs are stored next to the outputs
so that nothing depends on RNG reproducibility across torch versions.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from simple_pose_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def edge_joints():
    """Hand-built encoder vectors (SURVEY.md section 8c iii), shape (48, 64)."""
    rows = [
        (10.3, 20.7, 1.0),      # generic sub-pixel centre
        (53.9, 10.0, 1.0),      # ul_x = int(47.9) = 47 < 48 -> kept
        (54.1, 10.0, 1.0),      # ul_x = 48 >= W -> culled
        (-7.2, 5.0, 1.0),       # br_x = int(-0.2) = 0 -> kept (truncation toward zero)
        (-6.999, 10.0, 1.0),
        (-7.0, 10.0, 1.0),
        (-8.0, 10.0, 1.0),      # br_x = int(-1.0) = -1 -> culled
        (47.99, 63.99, 1.0),    # far corner
        (12.0, 30.0, 0.0),      # invisible -> zero map, weight 0
        (24.0, 70.1, 1.0),      # ul_y = 64 >= H -> culled
        (24.0, 69.9, 1.0),      # kept, centre below the map
        (20.0, 20.0, 2.0),      # vis = 2 (COCO "visible"): drawn, weight stays 2
        (20.0, 20.0, 0.5),      # weight 0.5 is not > 0.5 -> zero map, weight stays 0.5
        (0.0, 0.0, 1.0),
        (47.0, 63.0, 1.0),
        (23.5, 31.5, 1.0),
        (-3.25, -2.75, 1.0),
    ]
    return np.array(rows, dtype=np.float32)


def edge_maps(h=64, w=48):
    """Hand-built decoder vectors (SURVEY.md section 8c ii), one person, K = 17."""
    hm = np.zeros((1, 17, h, w), dtype=np.float32)
    hm[0, 0] = -1.0
    hm[0, 0, 0, 0] = -0.5                       # all negative -> (0,0), max -0.5
    # joint 1: all zero -> (0,0), 0.0
    hm[0, 2, 1, 1] = 1.0                        # peak at x=1: not refined
    hm[0, 3, 30, 20] = 1.0
    hm[0, 3, 30, 21] = 0.9                      # two-pixel peak -> refined in x only
    hm[0, 4, 10, 7] = 0.7
    hm[0, 4, 40, 30] = 0.7                      # tie -> lowest flat index
    hm[0, 5, 2, 2] = 1.0                        # smallest refinable coordinates
    hm[0, 6, h - 3, w - 3] = 1.0                # largest refinable coordinates
    hm[0, 7, h - 2, w - 2] = 1.0                # just outside
    yy, xx = np.mgrid[0:h, 0:w]
    hm[0, 8] = np.exp(-((xx - 20.3) ** 2 + (yy - 33.6) ** 2) / 8.0).astype(np.float32)
    # joint 9: positive peak in a negative sea -> blurred neighbourhood hits the 1e-10 clamp
    hm[0, 9] = -0.05
    hm[0, 9, 25, 25] = 0.3
    # joint 10: narrow positive blob on a negative plateau (partially clamped stencil)
    hm[0, 10] = (np.exp(-((xx - 12.4) ** 2 + (yy - 50.2) ** 2) / 2.0) - 0.02).astype(np.float32)
    # joint 11: peak near the left edge so the blur window is cut by the zero padding
    hm[0, 11] = np.exp(-((xx - 2.6) ** 2 + (yy - 3.4) ** 2) / 8.0).astype(np.float32)
    # joint 12: large amplitude
    hm[0, 12] = (37.5 * np.exp(-((xx - 40.2) ** 2 + (yy - 8.8) ** 2) / 8.0)).astype(np.float32)
    # joint 13: tiny amplitude
    hm[0, 13] = (1e-6 * np.exp(-((xx - 30.7) ** 2 + (yy - 20.1) ** 2) / 8.0)).astype(np.float32)
    # joint 14: flat plateau of ones (stencil all equal after blur? no, but symmetric)
    hm[0, 14, 20:30, 15:25] = 1.0
    # joint 15: peak on the last row
    hm[0, 15, h - 1, 10] = 0.8
    # joint 16: constant positive map -> argmax 0
    hm[0, 16] = 0.25
    return hm


def affine_fixtures(ref):
    """Box -> (centre, scale, area, trans_inv) from the reference's own box_to_center_scale +
    get_affine_transform (the body of BasicTransform.__call__, datasets/naive_data.py:44-56)."""
    out = {}
    for tag, inp, outp in (("a", (192, 256), (48, 64)), ("b", (288, 384), (72, 96))):
        boxes = synth.detection_boxes(96, seed=61, ratio_exact_every=8, ratio=inp[0] / inp[1]).numpy()
        boxes[1] = [10.0, 20.0, 10.0 + 1e-3, 20.0 + 2e-3]           # tiny box
        boxes[2] = [-1.5, 7.0, -0.5, 9.0]                           # centre x == -1: scale_mult skipped
        boxes[3] = [100.0, 50.0, 100.0, 50.0]                       # degenerate: singular system
        c, s, a, t64, f64 = [], [], [], [], []
        for x1, y1, x2, y2 in boxes.tolist():
            ci, si = ref.box_to_center_scale(x1, y1, x2 - x1, y2 - y1, inp[0] / inp[1])
            tf, ti = ref.get_affine_transform(ci, si, 0, outp)
            c.append(ci); s.append(si); a.append(si[0] * si[1]); t64.append(ti); f64.append(tf)
        out.update({"boxes_" + tag: boxes, "center_" + tag: np.stack(c), "scale_" + tag: np.stack(s),
                    "area_" + tag: np.array(a, dtype=np.float32), "tinv64_" + tag: np.stack(t64), "fwd64_" + tag: np.stack(f64),
                    "tinv_" + tag: torch.from_numpy(np.stack(t64)).float().numpy(),
                    "shapes_" + tag: np.array([inp, outp], dtype=np.int32)})
    np.savez_compressed(os.path.join(GOLDEN, "affine.npz"), **out)


def train_geometry_fixtures(ref):
    """Train-side caller of the encoder: the reference's own ``RefineSimpleTransform.__call__``
    (commons/transforms.py:193-223) with scripted draws (``ref_loader.run_train_transform``)."""
    out = {}
    for tag, inp, outp, n, keep_maps in (("a", (192, 256), (48, 64), 48, 4), ("b", (288, 384), (72, 96), 16, 1)):
        smp = synth.train_samples(n, seed=71)
        smp["rot"][1] = 180.0; smp["rot"][2] = -90.0; smp["rot"][3] = 1e-9     # cos < 0, cos ~ 1e-17, tiny angle
        smp["joints"][4, :, 2] = 0.0                                             # a person with no visible joint
        tinv, jin, hmaps, masks, boxes_out = [], [], [], [], []
        for i in range(n):
            kp = ref_loader.run_train_transform(ref, smp["boxes"][i].tolist(), int(smp["img_w"][i]), 480,
                                                smp["joints"][i].numpy(), float(smp["scale_ratio"][i]),
                                                float(smp["rot"][i]), bool(smp["flip"][i]), PAIRS, inp, outp)
            tinv.append(kp.trans_inv); jin.append(kp.joints); masks.append(kp.mask)
            if i < keep_maps:                                                    # maps of the first few only (size)
                hmaps.append(kp.heat_map)
            boxes_out.append(np.array(kp.box, dtype=np.float32))
        out.update({"img_w_" + tag: smp["img_w"].numpy(), "boxes_" + tag: smp["boxes"].numpy(),
                    "joints_" + tag: smp["joints"].numpy(), "scale_ratio_" + tag: smp["scale_ratio"].numpy(),
                    "rot_" + tag: smp["rot"].numpy(), "flip_" + tag: smp["flip"].numpy(),
                    "tinv64_" + tag: np.stack(tinv), "joints_input_" + tag: np.stack(jin),
                    "heat_map_" + tag: np.stack(hmaps), "mask_" + tag: np.stack(masks),
                    "box_out_" + tag: np.stack(boxes_out), "shapes_" + tag: np.array([inp, outp], dtype=np.int32)})
    np.savez_compressed(os.path.join(GOLDEN, "train_geom.npz"), **out)


def dark_fixtures(ref, dec_out):
    """Outputs of the reference's DarkPoseOriginalKeyPointDecoder (image space) on the decode fixtures'
    inputs (stored in decode.npz) plus a map whose Taylor step leaves the map on the negative side."""
    dark = ref.DarkPoseOriginalKeyPointDecoder()
    out = {}
    for tag in ("a", "b", "e"):
        c, m = dark(torch.from_numpy(dec_out["hm_" + tag]).clone(), torch.from_numpy(dec_out["tinv_" + tag]))
        out["img_" + tag], out["max_" + tag] = c.numpy(), m.numpy()
    hm_n = negative_step_maps()
    c, m = dark(torch.from_numpy(hm_n).clone(), synth.identity_affines(1))
    out["hm_n"], out["hsp_n"], out["max_n"] = hm_n, c.numpy(), m.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "dark_original.npz"), **out)


def negative_step_maps(h=64, w=48, want=17):
    """One person whose Taylor steps land at NEGATIVE coordinates (GaussTaylor clamps them to 0,
    DarkPoseOriginal keeps them): seeded search over smooth two-blob maps with a spike two pixels from a
    border, keeping joints whose DarkPoseOriginal result moves < 2e-5 px under a 1e-7 relative
    perturbation of the map (i.e. well-conditioned ones). Takes about a minute."""
    from oracle import heatmap_oracle as O
    rng = np.random.default_rng(3)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    eye = torch.eye(2, 3)[None]
    keep = []
    while len(keep) < want:
        hm = np.zeros((1, 17, h, w), np.float32)
        for k in range(17):
            cx, cy, s, amp = rng.uniform(-6, 1), rng.uniform(8, 50), rng.uniform(2.5, 6), rng.uniform(0.4, 0.9)
            if k % 2 == 0:
                f = amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * s * s))
            else:
                f = amp * np.exp(-((ys - cx) ** 2 + (xs - min(cy, 40)) ** 2) / (2 * s * s))
            f = f.astype(np.float32)
            f += (rng.uniform(0.0, 0.3) * np.exp(-((xs - rng.uniform(0, 10)) ** 2 + (ys - rng.uniform(0, 60)) ** 2) /
                                                 (2 * rng.uniform(2, 5) ** 2))).astype(np.float32)
            py, px = (int(rng.integers(10, 50)), 2) if k % 2 == 0 else (2, int(rng.integers(10, 40)))
            f[py, px] = f.max() * 1.02 + 1e-3
            hm[0, k] = f
        c = O.dark_original_decode(torch.from_numpy(hm), eye)[0].numpy()[0]
        neg = np.nonzero((c < 0).any(axis=1) & (np.abs(c).max(axis=1) < 60))[0]
        if len(neg) == 0:
            continue
        pert = hm * (1 + 1e-7 * rng.standard_normal(hm.shape)).astype(np.float32)
        c2 = O.dark_original_decode(torch.from_numpy(pert), eye)[0].numpy()[0]
        for k in neg:
            if np.abs(c2[k] - c[k]).max() < 2e-5 and len(keep) < want:
                keep.append(hm[0, k].copy())
    return np.stack(keep)[None]


def round2_fixtures(ref):
    """Round-2 rows: the solvers' loss expression on float16 / bfloat16 predictions with a GradScaler-style upstream
    gradient (processors/dp_pose_hrnet_solver.py:111-120: the half x float32 product promotes to float32 exactly as under
    autocast, the gradient comes back in the prediction's dtype), the reference's own ``kps_to_dict_``
    (metrics/pose_metrics.py:172-179) and ``oks_nms`` on tied scores made explicit by ``oracle.detie_scores``. Written to
    a file of its own so that the round-1 fixtures stay byte-identical."""
    from oracle import heatmap_oracle as O
    out = {}
    jl = synth.joints(5, height=16, width=12, seed=61).numpy()
    tl, wl = zip(*[ref.get_heat_map(j, 2.0, (12, 16)) for j in jl])
    target = torch.from_numpy(np.stack(tl))
    mask = torch.from_numpy(np.stack(wl))
    mask[0, 0] = 2.0
    pred32 = synth.predictions_like(target, seed=62)
    for name, dtype, scale in (("f16", torch.float16, 65536.0), ("bf16", torch.bfloat16, 1024.0)):
        pred = pred32.to(dtype).requires_grad_(True)
        crit = torch.nn.MSELoss()
        loss = 0.5 * crit(pred.mul(mask[[..., None, None]]), target.mul(mask[[..., None, None]]))
        (loss * scale).backward()
        assert pred.grad.dtype == dtype and loss.dtype == torch.float32
        out["amp_%s_pred_bits" % name] = pred.detach().view(torch.int16).numpy()
        out["amp_%s_grad_bits" % name] = pred.grad.view(torch.int16).numpy()
        out["amp_%s_loss" % name] = np.float32(loss.item())
        out["amp_%s_scale" % name] = np.float32(scale)
    out["amp_target"], out["amp_mask"] = target.numpy(), mask.numpy()
    # kps_to_dict_
    g = torch.Generator().manual_seed(63)
    coords = torch.randn(11, 17, 2, generator=g) * 100
    conf = torch.rand(11, 17, 1, generator=g)
    records = []
    ref.kps_to_dict_(coords, conf, list(range(700, 711)), records)
    out["dict_coords"], out["dict_conf"] = coords.numpy(), conf.numpy()
    out["dict_score"] = np.array([r["score"] for r in records], dtype=np.float64)
    out["dict_keypoints"] = np.array([r["keypoints"] for r in records], dtype=np.float64)
    out["dict_image_id"] = np.array([r["image_id"] for r in records], dtype=np.int64)
    # oks_nms with ties, de-tied "higher index first"
    kps, _, area, _ = synth.nms_groups(1, mean_group=40.0, seed=64)
    kps, area = kps.numpy(), area.numpy()
    tied = np.random.RandomState(65).choice([0.2, 0.4, 0.6], size=kps.shape[0])
    detied = O.detie_scores(tied)
    out["tie_kps"], out["tie_area"], out["tie_scores"], out["tie_detied"] = kps, area, tied, detied
    out["tie_keep"] = np.array([int(i) for i in ref.oks_nms(kps, detied, area, 0.9)], dtype=np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "round2.npz"), **out)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_loader.load()
    torch.manual_seed(0)
    affine_fixtures(ref)
    train_geometry_fixtures(ref)
    if "--only-affine" in sys.argv:
        return
    if "--only-dark" in sys.argv:
        dec = np.load(os.path.join(GOLDEN, "decode.npz"))
        dark_fixtures(ref, dec)
        return

    # ---------------------------------------------------------------- encode
    ja = synth.joints(3, height=64, width=48, seed=3).numpy()
    ta, wa = zip(*[ref.get_heat_map(j, 2.0, (48, 64)) for j in ja])
    jb = synth.joints(1, height=96, width=72, seed=4).numpy()
    tb, wb = zip(*[ref.get_heat_map(j, 2.0, (72, 96)) for j in jb])
    je = edge_joints()[None]
    te, we = ref.get_heat_map(je[0], 2.0, (48, 64))
    js = synth.joints(1, height=64, width=48, seed=5).numpy()       # other sigma
    ts, ws = ref.get_heat_map(js[0], 1.5, (48, 64))
    np.savez_compressed(
        os.path.join(GOLDEN, "encode.npz"),
        joints_a=ja, targets_a=np.stack(ta), weights_a=np.stack(wa),
        joints_b=jb, targets_b=np.stack(tb), weights_b=np.stack(wb),
        joints_e=je, targets_e=te[None], weights_e=we[None],
        joints_s=js, targets_s=ts[None], weights_s=ws[None])

    # ---------------------------------------------------------------- decode
    dec = ref.GaussTaylorKeyPointDecoder()
    np.savez_compressed(os.path.join(GOLDEN, "blur_weights.npz"),
                        w11=dec.blur_weights[0, 0].numpy())

    def run_decoder(hm, tinv):
        hm_t = torch.from_numpy(hm).clone()
        img, mx = dec(hm_t, torch.from_numpy(tinv))
        eye = synth.identity_affines(hm.shape[0]).numpy()
        hsp, _ = dec(torch.from_numpy(hm).clone(), torch.from_numpy(eye))
        b, k, h, w = hm.shape
        idx = torch.from_numpy(hm).reshape(b, k, -1).max(dim=-1)[1]
        return img.numpy(), hsp.numpy(), mx.numpy(), idx.numpy().astype(np.int32)

    out = {}
    hm_a = synth.heatmaps(3, height=64, width=48, seed=11, noise=0.01).numpy()
    ti_a = synth.inverse_affines(3, height=64, width=48, seed=11)[0].numpy()
    hm_b = synth.heatmaps(1, height=96, width=72, seed=12, noise=0.01).numpy()
    ti_b = synth.inverse_affines(1, height=96, width=72, seed=12)[0].numpy()
    hm_e = edge_maps()
    ti_e = synth.inverse_affines(1, height=64, width=48, seed=13)[0].numpy()
    for tag, hm, ti in (("a", hm_a, ti_a), ("b", hm_b, ti_b), ("e", hm_e, ti_e)):
        img, hsp, mx, idx = run_decoder(hm, ti)
        out.update({"hm_" + tag: hm, "tinv_" + tag: ti, "img_" + tag: img,
                    "hsp_" + tag: hsp, "max_" + tag: mx, "idx_" + tag: idx})
    # second opinion: the reference's NumPy/OpenCV decoder on a clone (it mutates its input)
    dark = ref.DarkPoseOriginalKeyPointDecoder()
    c2, _ = dark(torch.from_numpy(hm_a).clone(), torch.from_numpy(synth.identity_affines(3).numpy()))
    out["hsp_a_darkpose"] = c2.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "decode.npz"), **out)
    dark_fixtures(ref, out)

    # ---------------------------------------------------------------- flip-average + decode
    hm, hf = synth.flip_pair(2, height=64, width=48, seed=21, noise=0.01)
    perm = list(range(17))
    for a, b in PAIRS:
        perm[a], perm[b] = perm[b], perm[a]
    avg = 0.5 * (hm + hf.flip(-1)[:, perm])
    ti = synth.inverse_affines(2, seed=21)[0]
    img, mx = dec(avg.clone(), ti)
    hsp, _ = dec(avg.clone(), synth.identity_affines(2))
    # consistency of the permutation with the reference's flip_joints on a probe skeleton
    probe = np.stack([np.arange(17), np.arange(17), np.ones(17)], -1).astype(np.float32)
    _, flipped = ref.flip_joints(np.zeros((4, 48, 3), np.uint8), probe, PAIRS)
    assert [int(v) for v in flipped[:, 1]] == perm
    assert np.all(flipped[:, 0] == 48 - probe[perm, 0] - 1)
    np.savez_compressed(
        os.path.join(GOLDEN, "flip.npz"), hm=hm.numpy(), hm_flip=hf.numpy(), tinv=ti.numpy(),
        perm=np.array(perm, dtype=np.int32), img=img.numpy(), hsp=hsp.numpy(),
        max=mx.numpy(), idx=avg.reshape(2, 17, -1).max(dim=-1)[1].numpy().astype(np.int32))

    # ---------------------------------------------------------------- loss
    jl = synth.joints(4, height=16, width=12, seed=31).numpy()
    tl, wl = zip(*[ref.get_heat_map(j, 2.0, (12, 16)) for j in jl])
    target = torch.from_numpy(np.stack(tl))
    mask = torch.from_numpy(np.stack(wl))
    pred = synth.predictions_like(target, seed=32).requires_grad_(True)
    crit = torch.nn.MSELoss()
    loss = 0.5 * crit(pred.mul(mask[[..., None, None]]), target.mul(mask[[..., None, None]]))
    loss.backward()
    np.savez_compressed(os.path.join(GOLDEN, "loss.npz"), pred=pred.detach().numpy(),
                        target=target.numpy(), mask=mask.numpy(),
                        loss=np.float32(loss.item()), grad=pred.grad.numpy())

    # ---------------------------------------------------------------- OKS / NMS / rescoring
    kps, box, area, seg = synth.nms_groups(12, mean_group=14.0, seed=41)
    kps, box, area, seg = kps.numpy(), box.numpy(), area.numpy(), seg.numpy()
    scores = np.zeros_like(box)
    for i in range(kps.shape[0]):
        # eval.py:168-175
        kpt_scores = kps[i][:, -1]
        valid = kpt_scores > 0.2
        kpt_score = kpt_scores[valid].mean() if valid.sum() > 0 else 0.
        scores[i] = box[i] * kpt_score
    keep = np.zeros(kps.shape[0], dtype=np.uint8)
    picks = []
    for s in range(len(seg) - 1):
        lo, hi = seg[s], seg[s + 1]
        got = ref.oks_nms(kps[lo:hi], scores[lo:hi], area[lo:hi], 0.9)
        got = [int(g) + int(lo) for g in got]
        keep[got] = 1
        picks.extend(got)
    lo, hi = seg[0], seg[1]
    iou0 = ref.oks_iou(kps[lo], kps[lo:hi], area[lo], area[lo:hi])
    iou0_vis = ref.oks_iou(kps[lo], kps[lo:hi], area[lo], area[lo:hi], in_vis_thresh=0.5)
    np.savez_compressed(os.path.join(GOLDEN, "oks.npz"), kps=kps, box_scores=box, areas=area,
                        seg=seg, scores=scores, keep=keep, picks=np.array(picks, dtype=np.int32),
                        iou0=iou0, iou0_vis=iou0_vis)

    # ---------------------------------------------------------------- section 8f rows
    basic = ref.BasicKeyPointDecoder()
    bimg, bmax = basic(torch.from_numpy(hm_a).clone(), torch.from_numpy(ti_a))
    bhsp, _ = basic(torch.from_numpy(hm_a).clone(), synth.identity_affines(3))
    beimg, _ = basic(torch.from_numpy(hm_e).clone(), synth.identity_affines(1))
    jq = synth.joints(2, height=256, width=192, seed=51).numpy()
    tq, wq = zip(*[ref.get_heat_map_basic(j, 2.0, (48, 64), 4) for j in jq])
    acc = ref.HeatMapAcc()
    tgt = torch.from_numpy(np.stack(ta))
    prd = synth.predictions_like(tgt, seed=52, noise=0.2)
    msk = torch.from_numpy(np.stack(wa))[..., None, None]
    acc_val = acc(prd * msk, tgt * msk)
    np.savez_compressed(os.path.join(GOLDEN, "next_rows.npz"), basic_img=bimg.numpy(),
                        basic_hsp=bhsp.numpy(), basic_max=bmax.numpy(), basic_edge_hsp=beimg.numpy(),
                        joints_q=jq, targets_q=np.stack(tq), weights_q=np.stack(wq),
                        acc_pred=prd.numpy(), acc_value=np.float32(float(acc_val)))
    round2_fixtures(ref)
    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print("wrote fixtures to", GOLDEN, "(%.1f KB)" % (total / 1024))


if __name__ == "__main__":
    main()
