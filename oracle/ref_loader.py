"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (liangheming/simple_pose) from ``/root/reference``
so that the CPU restatement in ``oracle/heatmap_oracle.py`` can be pinned against it
and so that ``oracle/make_golden.py`` can freeze its outputs as fixtures.

``/root/reference`` exists only in the build container, never on the GPU box:
everything that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) uses the
committed fixtures in ``tests/golden/`` and the restatement instead.

Two shims are needed:

1. ``metrics/pose_metrics.py:6-7`` imports ``pycocotools`` at module top; it is not
   installed, so stub modules are registered in ``sys.modules`` first.
2. ``metrics/pose_metrics.py:102`` (``valid_mask[valid_mask] = derivative_valid_mask``)
   raises on torch >= 2.x because source and destination of the ``index_put_`` alias.
   The source text is patched in memory (one token: ``valid_mask.clone()`` as the
   index) and exec'd into a fresh module. Semantics are unchanged.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SIMPLE_POSE_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "metrics", "pose_metrics.py"))


def _stub_pycocotools():
    if "pycocotools" in sys.modules:
        return
    root = types.ModuleType("pycocotools")
    coco = types.ModuleType("pycocotools.coco")
    cocoeval = types.ModuleType("pycocotools.cocoeval")
    coco.COCO = type("COCO", (), {})
    cocoeval.COCOeval = type("COCOeval", (), {})
    root.coco, root.cocoeval = coco, cocoeval
    sys.modules["pycocotools"] = root
    sys.modules["pycocotools.coco"] = coco
    sys.modules["pycocotools.cocoeval"] = cocoeval


_cache = {}


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub_pycocotools()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    transforms = importlib.import_module("commons.transforms")
    joint_utils = importlib.import_module("commons.joint_utils")
    naive_data = importlib.import_module("datasets.naive_data")

    path = os.path.join(REFERENCE_ROOT, "metrics", "pose_metrics.py")
    with open(path, "r") as fh:
        text = fh.read()
    needle = "valid_mask[valid_mask] = derivative_valid_mask"
    assert text.count(needle) == 1, "reference decoder changed; re-check shim 2"
    text = text.replace(needle, "valid_mask[valid_mask.clone()] = derivative_valid_mask")
    pose_metrics = types.ModuleType("ref_pose_metrics_patched")
    pose_metrics.__file__ = path
    exec(compile(text, path, "exec"), pose_metrics.__dict__)

    ns = types.SimpleNamespace(
        transforms=transforms,
        joint_utils=joint_utils,
        naive_data=naive_data,
        pose_metrics=pose_metrics,
        get_heat_map=transforms.RefineSimpleTransform.get_heat_map,
        get_heat_map_basic=transforms.BasicSimpleTransform.get_heat_map,
        GaussTaylorKeyPointDecoder=pose_metrics.GaussTaylorKeyPointDecoder,
        DarkPoseOriginalKeyPointDecoder=pose_metrics.DarkPoseOriginalKeyPointDecoder,
        BasicKeyPointDecoder=pose_metrics.BasicKeyPointDecoder,
        HeatMapAcc=pose_metrics.HeatMapAcc,
        kps_to_dict_=pose_metrics.kps_to_dict_,
        oks_iou=naive_data.oks_iou,
        oks_iou_ori=naive_data.oks_iou_ori,
        oks_nms=naive_data.oks_nms,
        box_to_center_scale=joint_utils.box_to_center_scale,
        get_affine_transform=joint_utils.get_affine_transform,
        flip_joints=joint_utils.flip_joints,
    )
    _cache["ns"] = ns
    return ns


def run_train_transform(ns, box, img_w, img_h, joints, scale_ratio, rot, flip, joint_pairs,
                        input_shape=(192, 256), output_shape=(48, 64), basic=False):
    """The reference's own ``RefineSimpleTransform.__call__`` (commons/transforms.py:193-223) on one
    sample with its three ``np.random.uniform`` draws scripted (``rand_crop=False``: the box is taken
    as already cropped) and a blank image (the warp result is not looked at). Returns the mutated
    ``KeyPoints``: ``.trans_inv`` float64 [2,3], ``.joints`` (input-pixel joints), ``.heat_map``,
    ``.mask``, ``.box``."""
    import unittest.mock as mock
    import numpy as np
    T = ns.transforms
    cls = T.BasicSimpleTransform if basic else T.RefineSimpleTransform       # ``basic``: :118-148 instead of :193-223
    tr = cls(joint_pairs=[list(p) for p in joint_pairs], input_shape=tuple(input_shape),
             output_shape=tuple(output_shape), rand_crop=False)
    kp = T.KeyPoints("0.jpg", (int(img_w), int(img_h)), [float(v) for v in box], np.array(joints, dtype=np.float32))
    kp.img = np.zeros((int(img_h), int(img_w), 3), np.uint8)
    draws = iter([float(scale_ratio), float(rot), 0.0 if flip else 0.9])
    with mock.patch.object(np.random, "uniform", side_effect=lambda *a, **k: next(draws)):
        return tr(kp)


def run_eval_filter(kps, box_scores, areas, img_ids, in_vis_thre=0.2, oks_thre=0.9):
    """The reference's own ``temp_read_in_and_filter`` (eval.py:153-197): writes the per-person records
    the way ``predicts_by_pred`` does (eval.py:139-149: ``img_id``, ``score`` = box score, ``area``,
    ``kps`` = 51 floats) into a scratch directory, runs the function there with the pycocotools
    evaluation stubbed out, and returns the list of dicts it dumped (``image_id``, ``score``,
    ``keypoints``) in its output order."""
    import json
    import tempfile
    import numpy as np
    load()                                           # sys.path + pycocotools stub
    ev = importlib.import_module("eval")
    records = [{"img_id": int(i), "score": float(s), "area": float(a), "kps": np.asarray(k, dtype=np.float64).reshape(-1).tolist()}
               for k, s, a, i in zip(kps, box_scores, areas, img_ids)]
    keep_eval, cwd = ev.eval_kps, os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        try:
            os.chdir(tmp)
            with open("predicts_kps_temp.json", "w") as fh:
                json.dump(records, fh)
            ev.eval_kps = lambda *a, **k: None
            ev.temp_read_in_and_filter(in_vis_thre=in_vis_thre, oks_thre=oks_thre)
            with open("filter_kps_predicts.json") as fh:
                return json.load(fh)
        finally:
            ev.eval_kps = keep_eval
            os.chdir(cwd)
