"""CPU, build container only: every mirrored callable keeps the reference's signature -- same parameter
names in the same order with the same defaults (extra trailing keyword parameters are allowed), so the
call sites of the reference's solvers / eval script work unchanged."""
import inspect

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def _same_default(a, b):
    if a is inspect.Parameter.empty or b is inspect.Parameter.empty:
        return a is b
    try:
        import numpy as np
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):      # get_affine_transform(shift=array([0, 0]))
            return b is None or (isinstance(b, np.ndarray) and np.array_equal(a, b))
    except Exception:
        pass
    return a == b


def _check(ref_fn, our_fn, label):
    ref_p, our_p = _params(ref_fn), _params(our_fn)
    assert len(our_p) >= len(ref_p), (label, ref_p, our_p)
    for (rn, rd), (on, od) in zip(ref_p, our_p):
        assert rn == on, (label, rn, on)
        assert _same_default(rd, od), (label, rn, rd, od)
    for name, default in our_p[len(ref_p):]:
        assert default is not inspect.Parameter.empty, (label, "extra parameter without default", name)


def test_mirrored_signatures():
    ref = ref_loader.load()
    from simple_pose_b200.commons import joint_utils as ju, transforms as tr
    from simple_pose_b200.datasets import naive_data as nd
    from simple_pose_b200.metrics import pose_metrics as pm
    rt, rj, rn, rm = ref.transforms, ref.joint_utils, ref.naive_data, ref.pose_metrics
    pairs = [
        (rt.RefineSimpleTransform.__init__, tr.RefineSimpleTransform.__init__, "RefineSimpleTransform()"),
        (rt.RefineSimpleTransform.get_heat_map, tr.RefineSimpleTransform.get_heat_map, "Refine.get_heat_map"),
        (rt.BasicSimpleTransform.__init__, tr.BasicSimpleTransform.__init__, "BasicSimpleTransform()"),
        (rt.BasicSimpleTransform.get_heat_map, tr.BasicSimpleTransform.get_heat_map, "Basic.get_heat_map"),
        (rj.box_to_center_scale, ju.box_to_center_scale, "box_to_center_scale"),
        (rj.center_scale_to_box, ju.center_scale_to_box, "center_scale_to_box"),
        (rj.get_affine_transform, ju.get_affine_transform, "get_affine_transform"),
        (rj.affine_transform_batch, ju.affine_transform_batch, "affine_transform_batch"),
        (rj.flip_joints, ju.flip_joints, "flip_joints"),
        (rn.oks_iou, nd.oks_iou, "oks_iou"),
        (rn.oks_nms, nd.oks_nms, "oks_nms"),
        (rm.BasicKeyPointDecoder.heat_map_to_axis, pm.BasicKeyPointDecoder.heat_map_to_axis, "heat_map_to_axis"),
        (rm.BasicKeyPointDecoder.__call__, pm.BasicKeyPointDecoder.__call__, "BasicKeyPointDecoder.__call__"),
        (rm.GaussTaylorKeyPointDecoder.__init__, pm.GaussTaylorKeyPointDecoder.__init__, "GaussTaylorKeyPointDecoder()"),
        (rm.GaussTaylorKeyPointDecoder.__call__, pm.GaussTaylorKeyPointDecoder.__call__, "GaussTaylorKeyPointDecoder.__call__"),
        (rm.DarkPoseOriginalKeyPointDecoder.__init__, pm.DarkPoseOriginalKeyPointDecoder.__init__, "DarkPoseOriginalKeyPointDecoder()"),
        (rm.DarkPoseOriginalKeyPointDecoder.__call__, pm.DarkPoseOriginalKeyPointDecoder.__call__, "DarkPoseOriginal.__call__"),
        (rm.HeatMapAcc.__init__, pm.HeatMapAcc.__init__, "HeatMapAcc()"),
        (rm.HeatMapAcc.__call__, pm.HeatMapAcc.__call__, "HeatMapAcc.__call__"),
        (rm.kps_to_dict_, pm.kps_to_dict_, "kps_to_dict_"),
    ]
    for ref_fn, our_fn, label in pairs:
        _check(ref_fn, our_fn, label)
    # class relationships the reference's code relies on (HeatMapAcc calls heat_map_to_axis on the class)
    assert issubclass(pm.GaussTaylorKeyPointDecoder, pm.BasicKeyPointDecoder)
    assert isinstance(inspect.getattr_static(pm.BasicKeyPointDecoder, "heat_map_to_axis"), staticmethod)
    assert isinstance(inspect.getattr_static(tr.RefineSimpleTransform, "get_heat_map"), staticmethod)
