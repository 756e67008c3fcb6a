"""Mirror of the pieces of the reference's ``commons/joint_utils.py`` that sit either side of the
heatmap hot path: the left/right pair swap of ``flip_joints`` (:102-112) expressed as a channel
permutation (COCO's ``joint_pairs``, ``datasets/coco.py:26``, as the default), ``flip_joints`` and
``affine_transform_batch`` on the joints (:88-112), ``box_to_center_scale`` (:39-56) and
``get_affine_transform`` (:115-152, with rotation). All arithmetic runs in the sm_100a library.
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


def center_scale_to_box(center, scale):
    """``commons/joint_utils.py:59-69`` (host arithmetic on two 2-vectors, kept for drop-in completeness)."""
    w, h = scale[0], scale[1]
    xmin, ymin = center[0] - w * 0.5, center[1] - h * 0.5
    return (xmin, ymin, xmin + w, ymin + h)


def get_affine_transforms(center, scale, output_size, device=None, rot=None):
    """Batched ``get_affine_transform(center[i], scale[i], rot[i], output_size)``: center, scale [P,2]
    float32, rot [P] float64 degrees (None = 0) -> (trans [P,2,3], trans_inv [P,2,3]) float64 device
    tensors = the two ``cv.getAffineTransform`` results of the reference
    (``commons/joint_utils.py:149-150``); bit-identical for rot = 0."""
    import torch
    from .. import _abi
    c = _abi.to_device(center, torch.float32, device).reshape(-1, 2)
    s = _abi.to_device(scale, torch.float32, c.device).reshape(-1, 2)
    if s.shape[0] != c.shape[0]:
        raise ValueError("center and scale must both be [P, 2]")
    n = int(c.shape[0])
    r = None
    if rot is not None:
        r = _abi.to_device(rot, torch.float64, c.device).reshape(-1)
        if r.shape[0] != n:
            raise ValueError("rot must have one angle per (center, scale) pair")
    fwd = torch.empty((n, 2, 3), dtype=torch.float64, device=c.device)
    inv = torch.empty((n, 2, 3), dtype=torch.float64, device=c.device)
    with torch.cuda.device(c.device):
        _abi.check(_abi.lib().sp_center_scale_rot_affine_f64(c.data_ptr(), s.data_ptr(), _abi.ptr(r), None, inv.data_ptr(),
                                                             fwd.data_ptr(), n, int(output_size[0]), int(output_size[1]),
                                                             _abi.stream_ptr(c.device)))
    return fwd, inv


def get_affine_transform(center, scale, rot, output_size, shift=None):
    """Drop-in for ``commons/joint_utils.py:115-152`` with zero ``shift`` (every call site of the
    reference): ``(trans, trans_inv)`` as float64 [2,3] arrays."""
    import numpy as np
    if shift is not None and np.any(np.asarray(shift) != 0):
        raise NotImplementedError("simple_pose_b200: only shift = 0 (what every reference call site uses) runs on the device")
    scale = np.asarray(scale, dtype=np.float32).reshape(-1)
    if scale.size == 1:                                    # reference :129-130
        scale = np.array([scale[0], scale[0]], dtype=np.float32)
    fwd, inv = get_affine_transforms(np.asarray(center, dtype=np.float32).reshape(1, 2), scale.reshape(1, 2), output_size,
                                     rot=np.array([float(rot)], dtype=np.float64))
    return fwd[0].cpu().numpy(), inv[0].cpu().numpy()


def transform_joints(joints, trans=None, flip=None, img_w=None, joint_pairs=None):
    """Batched ``flip_joints`` (joint half) followed by ``affine_transform_batch``
    (``commons/joint_utils.py:88-112``): joints [P,K,3] float32, ``flip`` [P] bool + ``img_w`` [P] int
    (None = no flip), ``trans`` [P,2,3] float64 (None = no affine; rows with vis <= 0 are never
    mapped). Returns a new [P,K,3] float32 device tensor."""
    import torch
    from .. import _abi
    j = _abi.to_device(joints, torch.float32)
    if j.dim() != 3 or j.shape[-1] != 3:
        raise ValueError("joints must be [P, K, 3]")
    dev, n, k = j.device, int(j.shape[0]), int(j.shape[1])
    t = None if trans is None else _abi.to_device(trans, torch.float64, dev).reshape(n, 2, 3)
    fl = iw = perm = None
    if flip is not None:
        if img_w is None:
            raise ValueError("flip needs img_w")
        fl = _abi.to_device(flip, torch.uint8, dev).reshape(n)
        iw = _abi.to_device(img_w, torch.int32, dev).reshape(n)
        perm = torch.tensor(swap_permutation(k, joint_pairs), dtype=torch.int32, device=dev)
    out = torch.empty_like(j)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_transform_joints_f32(j.data_ptr(), _abi.ptr(t), _abi.ptr(fl), _abi.ptr(iw), _abi.ptr(perm),
                                                      out.data_ptr(), n, k, _abi.stream_ptr(dev)))
    return out


def affine_transform_batch(joints, t):
    """Drop-in for ``commons/joint_utils.py:88-99``: joints [K,3] ndarray, t [2,3] -> new [K,3] ndarray
    of the joints' dtype (the reference writes the float64 products back into the joint array)."""
    import numpy as np
    arr = np.asarray(joints)
    out = transform_joints(arr.astype(np.float32)[None], trans=np.asarray(t, dtype=np.float64)[None])
    return out[0].cpu().numpy().astype(arr.dtype, copy=False)


def flip_joints(img, joints_src, joint_pairs):
    """Drop-in for ``commons/joint_utils.py:102-112``: the image is mirrored on the host exactly as the
    reference does (a NumPy view), the joints on the device."""
    import numpy as np
    width = img.shape[1]
    arr = np.asarray(joints_src)
    out = transform_joints(arr.astype(np.float32)[None], flip=[True], img_w=[width], joint_pairs=joint_pairs)
    return np.fliplr(img), out[0].cpu().numpy().astype(arr.dtype, copy=False)
