"""Flip semantics shared by the encoder-side augmentation and the flip-test decoder.

Mirror of the two pieces of the reference's ``commons/joint_utils.py`` that the heatmap hot
path consumes: the left/right pair swap of ``flip_joints`` (:102-112) expressed as a channel
permutation, with COCO's ``joint_pairs`` (``datasets/coco.py:26``) as the default.
"""
COCO_JOINT_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def swap_permutation(num_joints, joint_pairs=None):
    """perm such that ``flipped[k] = original[perm[k]]`` after the pair swap."""
    pairs = COCO_JOINT_PAIRS if joint_pairs is None else joint_pairs
    perm = list(range(num_joints))
    for a, b in pairs:
        if not (0 <= a < num_joints and 0 <= b < num_joints):
            raise ValueError("joint pair (%d, %d) outside 0..%d" % (a, b, num_joints - 1))
        perm[a], perm[b] = perm[b], perm[a]
    return perm


def box_to_center_scale(x, y, w, h, aspect_ratio=1.0, scale_mult=1.25):
    """Drop-in for ``commons/joint_utils.py:39-56`` (one box -> float32 ``center`` [2], ``scale`` [2]).
    Runs ``sp_box_affine_f64`` on the current CUDA device; for whole detection files use
    ``datasets.naive_data.box_affines``."""
    from ..datasets.naive_data import box_affines
    out = box_affines([[x, y, w, h]], (aspect_ratio, 1.0), (48, 64), scale_mult, xywh=True)
    return out["center"][0].cpu().numpy(), out["scale"][0].cpu().numpy()


def get_affine_transforms(center, scale, output_size, device=None):
    """Batched ``get_affine_transform(center[i], scale[i], 0, output_size)``: center, scale [P,2]
    float32 -> (trans [P,2,3], trans_inv [P,2,3]) float64 device tensors, bit-identical to the two
    ``cv.getAffineTransform`` results of the reference (``commons/joint_utils.py:149-150``)."""
    import torch
    from .. import _abi
    c = _abi.to_device(center, torch.float32, device).reshape(-1, 2)
    s = _abi.to_device(scale, torch.float32, c.device).reshape(-1, 2)
    if s.shape[0] != c.shape[0]:
        raise ValueError("center and scale must both be [P, 2]")
    n = int(c.shape[0])
    fwd = torch.empty((n, 2, 3), dtype=torch.float64, device=c.device)
    inv = torch.empty((n, 2, 3), dtype=torch.float64, device=c.device)
    with torch.cuda.device(c.device):
        _abi.check(_abi.lib().sp_center_scale_affine_f64(c.data_ptr(), s.data_ptr(), None, inv.data_ptr(), fwd.data_ptr(), n,
                                                         int(output_size[0]), int(output_size[1]), _abi.stream_ptr(c.device)))
    return fwd, inv


def get_affine_transform(center, scale, rot, output_size, shift=None):
    """Drop-in for ``commons/joint_utils.py:115-152`` with ``rot == 0`` and zero ``shift`` (what the eval
    path uses, ``datasets/naive_data.py:50-51``): ``(trans, trans_inv)`` as float64 [2,3] arrays."""
    import numpy as np
    if rot != 0 or (shift is not None and np.any(np.asarray(shift) != 0)):
        raise NotImplementedError("simple_pose_b200: only rot = 0, shift = 0 (the eval-path transform) runs on the device")
    scale = np.asarray(scale, dtype=np.float32).reshape(-1)
    if scale.size == 1:                                    # reference :129-130
        scale = np.array([scale[0], scale[0]], dtype=np.float32)
    fwd, inv = get_affine_transforms(np.asarray(center, dtype=np.float32).reshape(1, 2), scale.reshape(1, 2), output_size)
    return fwd[0].cpu().numpy(), inv[0].cpu().numpy()
