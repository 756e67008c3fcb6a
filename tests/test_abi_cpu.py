"""CPU: the C-ABI library builds, loads without a GPU and exports exactly the symbols that
include/simple_pose_b200.h declares; host-side argument checks fire before any launch; the
product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from simple_pose_b200 import build, _abi
    build.build()
    return _abi.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "simple_pose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from simple_pose_b200 import _abi, build
    names = declared_symbols()
    assert len(names) >= 12
    assert sorted(_abi.SIGNATURES) == names            # ctypes table mirrors the header 1:1
    out = subprocess.run(["nm", "-D", "--defined-only", build.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sp_[a-z0-9_]+)", out))
    assert set(names) <= exported
    for n in names:
        assert getattr(lib, n) is not None


def test_header_compiles_as_c():
    src = '#include "simple_pose_b200.h"\nint main(void){return SP_ABI_VERSION-1;}\n'
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_library_is_sm100a_only(lib):
    from simple_pose_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_tma_bulk_copy_in_sass(lib):
    from simple_pose_b200 import build
    out = subprocess.run("cuobjdump -sass %s | grep -c UBLKCP" % build.lib_path(), shell=True, capture_output=True, text=True).stdout
    assert int(out.strip() or 0) >= 2                  # plain and flip variants of the decode kernel


def test_argument_errors_without_gpu(lib):
    from simple_pose_b200 import _abi
    assert lib.sp_abi_version() == _abi.ABI_VERSION == 2
    assert lib.sp_mse_workspace_bytes() >= 16
    assert b"aligned" in lib.sp_error_string(-2)
    assert lib.sp_error_string(0) == b"ok"
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.sp_encode_f32(None, None, None, 1, 17, 64, 48, 2.0, None) == -1
    assert lib.sp_encode_f32(16, 16, 16, 1, 17, 64, 48, -1.0, None) == -1
    assert lib.sp_decode_f32(None, None, None, None, None, None, None, None, 1, 17, 64, 48, 11, 0, None) == -1
    assert lib.sp_decode_f32(16, None, None, None, 16, 16, 16, None, 1, 17, 64, 48, 10, 0, None) == -4   # even ksize
    assert lib.sp_decode_f32(16, 16, None, None, 16, 16, 16, None, 1, 17, 64, 48, 11, 0, None) == -1     # flip w/o perm
    assert lib.sp_decode_f32(16, None, None, None, 16, 24, 16, None, 1, 17, 64, 48, 11, 0, None) == -2   # misaligned coords
    assert lib.sp_mse_fwd_bwd_f32(16, 16, 16, 16, 16, 16, 8, 1, 17, 3072, 1.0, 0, None) == -3             # workspace too small
    assert lib.sp_oks_nms_f64(16, 16, 16, 16, None, 16, 16, 4, 1, 16, 4, 0.9, 0, 0.0, None) == -1         # K != 17 w/o sigmas
    # the entry points added for the callers either side of the path and the workspace decode
    assert lib.sp_decode_workspace_bytes() == 16
    assert lib.sp_decode_ws_f32(16, None, None, None, 16, 16, 16, None, 1, 17, 64, 48, 11, 0, None, 16, None) == -1    # no workspace
    assert lib.sp_decode_ws_f32(16, None, None, None, 16, 16, 16, None, 1, 17, 64, 48, 11, 0, 16, 8, None) == -3       # too small
    assert lib.sp_decode_ws_f32(16, None, None, None, 16, 16, 16, None, 1, 17, 64, 48, 11, 0, 24, 16, None) == -2      # misaligned
    assert lib.sp_decode_f32(16, None, None, None, 16, 16, 16, None, 1, 17, 64, 48, 11, 4, None) == -1                  # unknown mode
    assert lib.sp_train_geometry_f32(None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                     1, 17, 192, 256, 48, 64, 1.25, None) == -1
    assert lib.sp_train_geometry_f32(16, None, 16, None, None, 16, None, 16, None, None, None, None, None, None,
                                     1, 17, 192, 256, 48, 64, 1.25, None) == -1                                        # flip w/o img_w / perm
    assert lib.sp_train_geometry_f32(16, None, 16, None, None, None, None, 16, None, None, None, None, None, None,
                                     1, 17, 192, 0, 48, 64, 1.25, None) == -1
    assert lib.sp_train_geometry_f32(None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                     0, 17, 192, 256, 48, 64, 1.25, None) == 0
    assert lib.sp_transform_joints_f32(16, None, None, None, None, 16, 1, 17, None) == -1                              # in place
    assert lib.sp_transform_joints_f32(16, None, 16, None, None, 32, 1, 17, None) == -1                                # flip w/o img_w / perm
    assert lib.sp_center_scale_rot_affine_f64(16, 16, None, None, None, None, 1, 48, 64, None) == -1                   # no output
    # round 2: decoder -> result rows, fused rescoring + NMS on the rows, kps_to_dict_ table, tuning reload
    assert lib.sp_decode_rows_f32(16, None, None, None, 16, None, 54, None, None, 1, 17, 64, 48, 11, 0, 16, 16, None) == -1   # no rows
    assert lib.sp_decode_rows_f32(16, None, None, None, 16, 16, 50, None, None, 1, 17, 64, 48, 11, 0, 16, 16, None) == -1    # stride < 3K
    assert lib.sp_eval_rows_nms_f32(16, 53, 16, 16, None, 16, None, None, 4, 1, 17, 4, 0.2, 0.9, None) == -1                 # stride < 3K+3
    assert lib.sp_eval_rows_nms_f32(16, 54, 16, None, None, 16, None, None, 4, 1, 17, 4, 0.2, 0.9, None) == -1               # no areas
    assert lib.sp_eval_rows_nms_f32(None, 54, None, None, None, None, None, None, 0, 0, 17, 0, 0.2, 0.9, None) == 0
    assert lib.sp_person_rows_f32(None, 16, 16, 1, 17, None) == -1
    assert lib.sp_person_rows_f32(None, None, None, 0, 17, None) == 0
    assert lib.sp_reload_tuning() == 0
    # one-launch step, autocast-dtype loss, fan-out NMS: argument errors are caught before any CUDA call
    step_ok = [16, 16, 16, 16, 16, 16, 16, 16, 16, 16, None, None, 16, 1 << 16, 1, 17, 64, 48, 2.0, 11, 1.0, None]
    assert lib.sp_step_f32(*([None] + step_ok[1:])) == -1                                     # no joints
    assert lib.sp_step_f32(*(step_ok[:8] + [16, None] + step_ok[10:])) == -1                  # coords without maxval
    assert lib.sp_step_f32(*(step_ok[:10] + [16, None] + step_ok[12:])) == -1                 # pred_xy without label_xy
    assert lib.sp_step_f32(*(step_ok[:17] + [50] + step_ok[18:])) == -4                       # W % 4 != 0
    assert lib.sp_step_f32(*(step_ok[:19] + [9] + step_ok[20:])) == -4                        # only the 11 x 11 blur
    assert lib.sp_step_f32(*(step_ok[:13] + [8] + step_ok[14:])) == -3                        # workspace too small
    assert lib.sp_step_f32(*(step_ok[:1] + [24] + step_ok[2:])) == -2                         # misaligned pred
    assert lib.sp_mse_fwd_bwd(16, 7, 16, 16, 16, 16, 16, 1 << 16, 1, 17, 3072, 1.0, None, 0, None) == -1      # unknown dtype
    assert lib.sp_mse_fwd_bwd(16, 1, 16, 16, None, None, 16, 1 << 16, 1, 17, 3072, 1.0, None, 0, None) == -1  # nothing to compute
    assert lib.sp_mse_fwd_bwd(16, 1, 16, 16, 16, 16, 16, 8, 1, 17, 3072, 1.0, None, 0, None) == -3            # workspace too small
    assert lib.sp_encode_mse_fwd_bwd(16, 16, 9, 16, None, 16, 16, None, None, 16, 1 << 16, 1, 17, 64, 48, 2.0, 1.0, None, None) == -1
    fan = [16, 54, 16, 16, None, 16, None, None, 4, 1, 17, 4, 0.2, 0.9]
    assert lib.sp_eval_rows_nms_fanout_f32(*(fan + [None, None, 2, 0, 0, None])) == -1        # world 2 without any peer mapping
    assert lib.sp_eval_rows_nms_fanout_f32(*(fan + [16, None, 2, 2, 0, None])) == -1          # rank outside the world
    assert lib.sp_eval_rows_nms_fanout_f32(*(fan + [16, None, 2, 0, 7, None])) == -2          # odd slot offset: rows must stay 8-byte aligned
    assert lib.sp_eval_rows_nms_fanout_f32(*([16, 55] + fan[2:] + [16, None, 2, 0, 0, None])) == -2   # odd row stride
    # empty batches are a no-op success
    assert lib.sp_encode_f32(None, None, None, 0, 17, 64, 48, 2.0, None) == 0
    assert lib.sp_decode_f32(16, None, None, None, 16, 16, 16, None, 0, 17, 64, 48, 11, 0, None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from simple_pose_b200.commons.transforms import RefineSimpleTransform, encode_heat_maps
    from simple_pose_b200.metrics.pose_metrics import GaussTaylorKeyPointDecoder
    from simple_pose_b200.processors.loss import JointsMSELoss
    from simple_pose_b200.datasets.naive_data import oks_nms
    with pytest.raises(RuntimeError, match="no CPU path"):
        RefineSimpleTransform.get_heat_map(np.zeros((17, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        encode_heat_maps(torch.zeros(2, 17, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        GaussTaylorKeyPointDecoder()(torch.zeros(1, 17, 64, 48), torch.zeros(1, 2, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        JointsMSELoss()(torch.zeros(1, 17, 8, 8), torch.zeros(1, 17, 8, 8), torch.ones(1, 17))
    with pytest.raises(RuntimeError, match="no CPU path"):
        oks_nms(np.zeros((2, 17, 3)), np.array([.5, .4]), np.ones(2), 0.9)
    from simple_pose_b200.commons.transforms import train_geometry, train_targets, BasicSimpleTransform
    from simple_pose_b200.commons import joint_utils as ju
    from simple_pose_b200.metrics.pose_metrics import DarkPoseOriginalKeyPointDecoder
    with pytest.raises(RuntimeError, match="no CPU path"):
        train_geometry(np.zeros((2, 4)), np.zeros((2, 17, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        train_targets(np.zeros((2, 4)), np.zeros((2, 17, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        BasicSimpleTransform().joint_targets(np.zeros((2, 4)), np.zeros((2, 17, 3), np.float32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ju.get_affine_transform(np.zeros(2, np.float32), np.ones(2, np.float32), 30.0, (48, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ju.flip_joints(np.zeros((4, 8, 3), np.uint8), np.zeros((17, 3), np.float32), [[1, 2]])
    with pytest.raises(RuntimeError, match="no CPU path"):
        ju.affine_transform_batch(np.zeros((17, 3), np.float32), np.zeros((2, 3)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DarkPoseOriginalKeyPointDecoder()(torch.zeros(1, 17, 64, 48), torch.zeros(1, 2, 3))
    assert ju.center_scale_to_box(np.array([10.0, 20.0], np.float32), np.array([4.0, 8.0], np.float32)) == (8.0, 16.0, 12.0, 24.0)


def test_missing_extension_is_loud(monkeypatch, tmp_path):
    from simple_pose_b200 import _abi
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setenv("SIMPLE_POSE_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_abi.ExtensionMissing, match="no CPU fallback"):
        _abi.lib()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under simple_pose_b200/ may reference it."""
    pkg = os.path.join(ROOT, "simple_pose_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(base, f)
                assert "/root/reference" not in text, os.path.join(base, f)


def test_swap_permutation():
    from simple_pose_b200.commons.joint_utils import swap_permutation
    assert swap_permutation(17) == [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    assert swap_permutation(4, [[0, 3]]) == [3, 1, 2, 0]
    with pytest.raises(ValueError):
        swap_permutation(4, [[0, 9]])


class _PerSampleEncodeDataset(torch.utils.data.Dataset):
    """What MSCOCO.__getitem__ does with the transform (reference datasets/coco.py:58-60)."""

    def __len__(self):
        return 4

    def __getitem__(self, i):
        import numpy as np
        from simple_pose_b200.commons.transforms import RefineSimpleTransform
        joints = np.array([[10.0, 20.0, 1.0]] * 17, dtype=np.float32)
        targets, weights = RefineSimpleTransform.get_heat_map(joints)
        return torch.from_numpy(targets), torch.from_numpy(weights)


def test_per_sample_encoder_refuses_dataloader_workers():
    """The per-sample drop-in cannot create a CUDA context in a forked DataLoader worker: it must say so (and
    point at the batched path) instead of failing inside CUDA."""
    loader = torch.utils.data.DataLoader(_PerSampleEncodeDataset(), batch_size=2, num_workers=1)
    with pytest.raises(RuntimeError, match="DataLoader worker"):
        next(iter(loader))
