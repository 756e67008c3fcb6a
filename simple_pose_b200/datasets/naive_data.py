"""OKS similarity, keypoint NMS and eval rescoring: drop-ins for ``oks_iou`` / ``oks_nms`` of the
reference's ``datasets/naive_data.py:120-173`` and the per-image loop of ``eval.py:153-197``.

``oks_nms_batched`` is the form the hardware wants: all images of an eval run in one launch
(one CTA per image), float64 like the reference. ``oks_iou`` / ``oks_nms`` keep the reference's
per-image NumPy signatures on top of it.
"""
import numpy as np
import torch

from .. import _abi


def _f64(x, device=None):
    return _abi.to_device(x, torch.float64, device)


def oks_iou(pick_kps, candi_kps, pick_area, candi_area, sigmas=None, in_vis_thresh=None):
    """pick_kps [K,3], candi_kps [n,K,3], areas -> ndarray [n] float64 (reference :120-150)."""
    pk = _f64(np.asarray(pick_kps, dtype=np.float64))
    ck = _f64(np.asarray(candi_kps, dtype=np.float64), pk.device)
    n, k = int(ck.shape[0]), int(ck.shape[1])
    pa = _f64(np.asarray(pick_area, dtype=np.float64).reshape(-1)[:1], pk.device)
    ca = _f64(np.asarray(candi_area, dtype=np.float64).reshape(-1), pk.device)
    sg = _f64(np.asarray(sigmas, dtype=np.float64), pk.device) if isinstance(sigmas, np.ndarray) else None
    out = torch.empty(n, dtype=torch.float64, device=pk.device)
    with torch.cuda.device(pk.device):
        _abi.check(_abi.lib().sp_oks_iou_f64(pk.data_ptr(), ck.data_ptr(), pa.data_ptr(), ca.data_ptr(),
                                             _abi.ptr(sg), out.data_ptr(), n, k,
                                             0 if in_vis_thresh is None else 1,
                                             0.0 if in_vis_thresh is None else float(in_vis_thresh),
                                             _abi.stream_ptr(pk.device)))
    return out.cpu().numpy()


def oks_nms_batched(kps, scores, areas, seg_offsets, thresh=0.9, sigmas=None, in_vis_thresh=None,
                    max_per_image=None):
    """Segmented greedy OKS-NMS. kps [N,K,3], scores [N], areas [N] (float64, device or host),
    seg_offsets [I+1] int32 (host or device). Returns (keep [N] uint8, rank [N] int32) on the
    device; ``rank`` is each person's position in its image's descending-score order."""
    k_t = _f64(kps)
    dev = k_t.device
    s_t, a_t = _f64(scores, dev), _f64(areas, dev)
    n, k = int(k_t.shape[0]), int(k_t.shape[1])
    if isinstance(seg_offsets, torch.Tensor) and seg_offsets.is_cuda:
        seg_dev = _abi.dense(seg_offsets, torch.int32)
        if max_per_image is None:
            max_per_image = int((seg_dev[1:] - seg_dev[:-1]).max().item()) if seg_dev.numel() > 1 else 0
    else:
        seg_host = np.ascontiguousarray(np.asarray(seg_offsets, dtype=np.int32))
        if max_per_image is None:
            max_per_image = int(np.diff(seg_host).max()) if seg_host.size > 1 else 0
        seg_dev = torch.from_numpy(seg_host).to(dev)
    images = int(seg_dev.numel()) - 1
    sg = _f64(np.asarray(sigmas, dtype=np.float64), dev) if isinstance(sigmas, np.ndarray) else None
    keep = torch.zeros(n, dtype=torch.uint8, device=dev)
    rank = torch.zeros(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_oks_nms_f64(k_t.data_ptr(), s_t.data_ptr(), a_t.data_ptr(), seg_dev.data_ptr(),
                                             _abi.ptr(sg), keep.data_ptr(), rank.data_ptr(), n, images, k,
                                             int(max_per_image), float(thresh),
                                             0 if in_vis_thresh is None else 1,
                                             0.0 if in_vis_thresh is None else float(in_vis_thresh),
                                             _abi.stream_ptr(dev)))
    return keep, rank


def oks_nms(kps, scores, areas, thresh, sigmas=None, in_vis_thresh=None):
    """One image: returns the kept indices in pick order (reference :153-173)."""
    scores = np.asarray(scores, dtype=np.float64)
    n = int(scores.shape[0])
    if n == 0:
        return []
    keep, rank = oks_nms_batched(np.asarray(kps, dtype=np.float64), scores, np.asarray(areas, dtype=np.float64),
                                 np.array([0, n], dtype=np.int32), thresh, sigmas, in_vis_thresh, max_per_image=n)
    keep = keep.cpu().numpy().astype(bool)
    rank = rank.cpu().numpy()
    kept = np.nonzero(keep)[0]
    return [int(i) for i in kept[np.argsort(rank[kept], kind="stable")]]


def rescore(kps, box_scores, in_vis_thre=0.2):
    """eval.py:168-175 for all persons at once: box_score * mean(joint conf > thr) (0 if none).
    Returns a float64 device tensor [N]."""
    k_t = _f64(kps)
    b_t = _f64(box_scores, k_t.device)
    out = torch.empty_like(b_t)
    with torch.cuda.device(k_t.device):
        _abi.check(_abi.lib().sp_rescore_f64(k_t.data_ptr(), b_t.data_ptr(), out.data_ptr(), int(k_t.shape[0]),
                                             int(k_t.shape[1]), float(in_vis_thre), _abi.stream_ptr(k_t.device)))
    return out


def pack_keypoints(coords, max_val):
    """Decoder output (float32 [N,K,2], [N,K,1]) -> float64 [N,K,3] (x, y, conf) on the device,
    the array ``oks_nms`` consumes (eval.py:138,166 without the JSON round trip)."""
    dev = _abi.require_cuda(coords, max_val)
    c = _abi.dense(coords, torch.float32)
    m = _abi.dense(max_val, torch.float32)
    n, k = int(c.shape[0]), int(c.shape[1])
    out = torch.empty((n, k, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_pack_kps_f64(c.data_ptr(), m.data_ptr(), out.data_ptr(), n, k, _abi.stream_ptr(dev)))
    return out


def rescore_and_nms(kps, box_scores, areas, seg_offsets, in_vis_thre=0.2, oks_thre=0.9, max_per_image=None):
    """eval.py:153-197 on the device: rescoring, then per-image OKS-NMS. Returns
    (keep uint8 [N], scores float64 [N], rank int32 [N]) device tensors."""
    k_t = _f64(kps)
    scores = rescore(k_t, box_scores, in_vis_thre)
    keep, rank = oks_nms_batched(k_t, scores, areas, seg_offsets, oks_thre, max_per_image=max_per_image)
    return keep, scores, rank


def box_affines(boxes_xyxy, input_shape=(192, 256), output_shape=(48, 64), scale_mult=1.25, want_f64=False,
                xywh=False):
    """``BasicTransform.__call__`` without the image warp (reference ``datasets/naive_data.py:44-56``)
    for all detection boxes at once: boxes [P,4] (x1, y1, x2, y2; host list/array or device tensor,
    float64) -> dict of device tensors ``center`` [P,2], ``scale`` [P,2], ``area`` [P] (float32) and
    ``trans_inv`` [P,2,3] float32, the heatmap -> image affine the decoder consumes (bit-identical to
    ``get_affine_transform(center, scale, 0, output_shape)[1]`` after ``collate_fn``'s ``.float()``).
    ``xywh=True`` reads the rows as (x, y, w, h), the arguments of ``box_to_center_scale``.
    ``want_f64`` adds ``trans_inv_f64``, the unrounded matrix, and ``trans_f64``, the forward
    (image -> heatmap) matrix -- the two float64 return values of ``get_affine_transform``."""
    b = _f64(np.asarray(boxes_xyxy, dtype=np.float64) if not isinstance(boxes_xyxy, torch.Tensor) else boxes_xyxy)
    if b.dim() != 2 or b.shape[1] != 4:
        raise ValueError("boxes must be [P, 4]: (x1, y1, x2, y2), or (x, y, w, h) with xywh=True")
    dev, n = b.device, int(b.shape[0])
    out = {"center": torch.empty((n, 2), dtype=torch.float32, device=dev),
           "scale": torch.empty((n, 2), dtype=torch.float32, device=dev),
           "area": torch.empty((n,), dtype=torch.float32, device=dev),
           "trans_inv": torch.empty((n, 2, 3), dtype=torch.float32, device=dev)}
    if want_f64:
        out["trans_inv_f64"] = torch.empty((n, 2, 3), dtype=torch.float64, device=dev)
        out["trans_f64"] = torch.empty((n, 2, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_box_affine_f64(
            b.data_ptr(), _abi.SP_BOX_XYWH if xywh else _abi.SP_BOX_XYXY, out["center"].data_ptr(), out["scale"].data_ptr(), out["area"].data_ptr(),
            out["trans_inv"].data_ptr(), _abi.ptr(out.get("trans_inv_f64")), _abi.ptr(out.get("trans_f64")), n,
            float(input_shape[0]) / float(input_shape[1]), int(output_shape[0]), int(output_shape[1]),
            float(scale_mult), _abi.stream_ptr(dev)))
    return out
