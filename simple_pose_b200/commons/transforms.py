"""Target encoder: drop-in for ``RefineSimpleTransform.get_heat_map`` of the reference
(``commons/transforms.py:167-191``), computed by the sm_100a kernel ``sp_encode_f32``.

* ``encode_heat_maps(joints[B,K,3], sigma, shape)`` is the batched device form: what
  ``MSCOCO.collate_fn`` (``datasets/coco.py:124-148``) stacks out of per-sample results, made in
  one launch on the training device from 204 bytes of joints per person instead of shipping
  209 KB of float32 maps per person over PCIe.
* ``RefineSimpleTransform.get_heat_map(joints, sigma, shape)`` keeps the reference's
  per-sample NumPy signature (B = 1 wrapper).
"""
import numpy as np
import torch

from .. import _abi


def encode_heat_maps(joints, sigma=2.0, shape=(48, 64), out=None):
    """joints [B,K,3] (x, y, vis) in heatmap pixels; ``shape`` is (W, H) as in the reference.

    Returns (targets [B,K,H,W] float32, weights [B,K] float32) on the CUDA device. Host inputs
    (NumPy / CPU tensors) are copied to the current CUDA device first."""
    j = _abi.to_device(joints, torch.float32)
    if j.dim() != 3 or j.shape[-1] != 3:
        raise ValueError("joints must be [B, K, 3], got %s" % (tuple(j.shape),))
    width, height = int(shape[0]), int(shape[1])
    b, k = int(j.shape[0]), int(j.shape[1])
    dev = j.device
    if out is None:
        targets = torch.empty((b, k, height, width), dtype=torch.float32, device=dev)
        weights = torch.empty((b, k), dtype=torch.float32, device=dev)
    else:
        targets, weights = out
        _abi.require_cuda(j, targets, weights)
        if tuple(targets.shape) != (b, k, height, width) or tuple(weights.shape) != (b, k) \
                or targets.dtype != torch.float32 or weights.dtype != torch.float32 \
                or not targets.is_contiguous() or not weights.is_contiguous():
            raise ValueError("out buffers have the wrong shape/dtype/layout")
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_encode_f32(j.data_ptr(), targets.data_ptr(), weights.data_ptr(),
                                            b, k, height, width, float(sigma), _abi.stream_ptr(dev)))
    return targets, weights


def train_geometry(boxes, joints, img_w=None, scale_ratio=None, rot=None, flip=None, joint_pairs=None,
                   input_shape=(192, 256), output_shape=(48, 64), scale_mult=1.25, want_input=False):
    """Joint/affine half of ``RefineSimpleTransform.__call__`` (reference ``commons/transforms.py:193-223``)
    for a whole batch, on the device, with the random draws passed in (``box_crop`` and the image warp
    stay with the CPU loader): boxes [P,4] float64 (x1, y1, x2, y2) after the crop, joints [P,K,3]
    float32 image pixels, ``scale_ratio`` / ``rot`` [P] float64 (draws of :202,:204; None = 1 / 0),
    ``flip`` [P] bool + ``img_w`` [P] int (draw of :208; None = no flip).

    Returns a dict of device tensors: ``joints_hm`` [P,K,3] (the encoder's input, :218),
    ``trans_inv`` [P,2,3] float32 (``joint_info.trans_inv`` after ``collate_fn``'s ``.float()``),
    ``trans_inv_f64``, ``center`` / ``scale`` [P,2] float32 after augmentation and, with
    ``want_input``, ``joints_input`` (``joint_info.joints``, :217) and ``img_trans_f64`` (the matrix
    of the image warp, :212)."""
    from .joint_utils import swap_permutation
    b = _abi.to_device(np.asarray(boxes, dtype=np.float64) if not isinstance(boxes, torch.Tensor) else boxes, torch.float64)
    dev = b.device
    j = _abi.to_device(joints, torch.float32, dev)
    if b.dim() != 2 or b.shape[1] != 4 or j.dim() != 3 or j.shape[-1] != 3 or j.shape[0] != b.shape[0]:
        raise ValueError("boxes must be [P, 4] and joints [P, K, 3]")
    n, k = int(j.shape[0]), int(j.shape[1])

    def opt(x, dtype):
        if x is None:
            return None
        t = _abi.to_device(x, dtype, dev).reshape(-1)
        if t.shape[0] != n:
            raise ValueError("per-person arguments must have %d entries" % n)
        return t
    sr, rt = opt(scale_ratio, torch.float64), opt(rot, torch.float64)
    fl, iw = opt(flip, torch.uint8), opt(img_w, torch.int32)
    perm = None
    if fl is not None:
        if iw is None:
            raise ValueError("flip needs img_w")
        perm = torch.tensor(swap_permutation(k, joint_pairs), dtype=torch.int32, device=dev)
    out = {"joints_hm": torch.empty_like(j),
           "trans_inv": torch.empty((n, 2, 3), dtype=torch.float32, device=dev),
           "trans_inv_f64": torch.empty((n, 2, 3), dtype=torch.float64, device=dev),
           "center": torch.empty((n, 2), dtype=torch.float32, device=dev),
           "scale": torch.empty((n, 2), dtype=torch.float32, device=dev)}
    if want_input:
        out["joints_input"] = torch.empty_like(j)
        out["img_trans_f64"] = torch.empty((n, 2, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_train_geometry_f32(
            b.data_ptr(), _abi.ptr(iw), j.data_ptr(), _abi.ptr(sr), _abi.ptr(rt), _abi.ptr(fl), _abi.ptr(perm),
            out["joints_hm"].data_ptr(), _abi.ptr(out.get("joints_input")), out["trans_inv"].data_ptr(),
            out["trans_inv_f64"].data_ptr(), _abi.ptr(out.get("img_trans_f64")), out["center"].data_ptr(),
            out["scale"].data_ptr(), n, k, int(input_shape[0]), int(input_shape[1]), int(output_shape[0]),
            int(output_shape[1]), float(scale_mult), _abi.stream_ptr(dev)))
    return out


def train_targets(boxes, joints, img_w=None, scale_ratio=None, rot=None, flip=None, joint_pairs=None,
                  input_shape=(192, 256), output_shape=(48, 64), sigma=2.0):
    """What ``MSCOCO.collate_fn`` (reference ``datasets/coco.py:124-148``) stacks out of
    ``RefineSimpleTransform.__call__`` results, minus the images: ``(heat_maps [P,K,H,W], masks [P,K],
    trans_invs [P,2,3])`` float32 on the device, from boxes + image-pixel joints + the augmentation
    draws (``train_geometry`` then ``encode_heat_maps``: two launches, 240 B in per person)."""
    geo = train_geometry(boxes, joints, img_w, scale_ratio, rot, flip, joint_pairs, input_shape, output_shape)
    heat_maps, masks = encode_heat_maps(geo["joints_hm"], sigma, output_shape)
    return heat_maps, masks, geo["trans_inv"]


class RefineSimpleTransform(object):
    """The hot-path members of the reference class: ``get_heat_map`` and the joint/affine half of
    ``__call__`` (``joint_targets``); ``box_crop`` and the image warp stay with the reference's CPU
    loader."""

    def __init__(self, joint_pairs=None, input_shape=(192, 256), output_shape=(48, 64),
                 scale=(0.7, 1.3), ratio=(-40, 40), rand_crop=True):
        self.input_shape = input_shape
        self.output_shape = output_shape
        self.joint_pairs = joint_pairs
        self.w_h_ratio = self.input_shape[0] / self.input_shape[1]
        self.scale = scale
        self.ratio = ratio
        self.rand_crop = rand_crop

    @staticmethod
    def get_heat_map(joints, sigma=2.0, shape=(48, 64)):
        """joints [K,3] ndarray -> (targets [K,H,W] float32 ndarray, weights [K] float32 ndarray).
        One person per call costs an H2D copy, a launch, a D2H copy and a sync; it exists for signature parity
        and small scripts. It cannot run in DataLoader worker processes (it raises there): the training path is
        the batched ``encode_heat_maps`` / ``EncodeJointsMSELoss`` on the device."""
        _abi.refuse_loader_worker("RefineSimpleTransform.get_heat_map")
        arr = np.ascontiguousarray(np.asarray(joints, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("joints must be [K, 3]")
        targets, weights = encode_heat_maps(arr[None], sigma, shape)
        return targets[0].cpu().numpy(), weights[0].cpu().numpy()

    def joint_targets(self, boxes, joints, img_w=None, scale_ratio=None, rot=None, flip=None, sigma=2.0):
        """Batched ``__call__`` (reference :193-223) without ``box_crop`` and the image warp: returns
        ``(heat_maps, masks, trans_invs)`` device tensors for this transform's shapes and joint pairs.
        ``flip`` is ignored when the transform has no ``joint_pairs`` (reference :207)."""
        if self.joint_pairs is None:
            flip = None
        return train_targets(boxes, joints, img_w, scale_ratio, rot, flip, self.joint_pairs, self.input_shape,
                             self.output_shape, sigma)


def basic_gaussian_table(sigma=2.0):
    """The fixed patch of ``BasicSimpleTransform.get_heat_map`` (reference :93-99), computed with
    the same NumPy float32 expressions so that its bits are the reference's."""
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    return np.ascontiguousarray(np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * (sigma ** 2))), dtype=np.float32)


_table_cache = {}


def encode_heat_maps_basic(joints, sigma=2.0, shape=(48, 64), stride=4, out=None):
    """Batched ``BasicSimpleTransform.get_heat_map``: joints [B,K,3] in INPUT pixels ->
    (targets [B,K,H,W] float32, weights [B,K] float32) on the CUDA device."""
    j = _abi.to_device(joints, torch.float32)
    if j.dim() != 3 or j.shape[-1] != 3:
        raise ValueError("joints must be [B, K, 3], got %s" % (tuple(j.shape),))
    width, height = int(shape[0]), int(shape[1])
    b, k = int(j.shape[0]), int(j.shape[1])
    dev = j.device
    key = (dev, float(sigma))
    table = _table_cache.get(key)
    if table is None:
        table = torch.from_numpy(basic_gaussian_table(sigma)).to(dev)
        _table_cache[key] = table
    if out is None:
        targets = torch.empty((b, k, height, width), dtype=torch.float32, device=dev)
        weights = torch.empty((b, k), dtype=torch.float32, device=dev)
    else:
        targets, weights = out
        _abi.require_cuda(j, targets, weights)
        if tuple(targets.shape) != (b, k, height, width) or tuple(weights.shape) != (b, k) \
                or targets.dtype != torch.float32 or weights.dtype != torch.float32 \
                or not targets.is_contiguous() or not weights.is_contiguous():
            raise ValueError("out buffers have the wrong shape/dtype/layout")
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_encode_basic_f32(j.data_ptr(), table.data_ptr(), targets.data_ptr(), weights.data_ptr(),
                                                  b, k, height, width, float(sigma), int(stride), int(table.shape[0]),
                                                  _abi.stream_ptr(dev)))
    return targets, weights


def train_targets_basic(boxes, joints, img_w=None, scale_ratio=None, rot=None, flip=None, joint_pairs=None,
                        input_shape=(192, 256), output_shape=(48, 64), sigma=2.0, stride=4):
    """``BasicSimpleTransform.__call__`` (reference ``commons/transforms.py:118-148``) without the image work:
    as ``train_targets`` but the quantised 13x13 encoder runs on the INPUT-pixel joints (``:141-143``).
    Returns ``(heat_maps, masks, trans_invs)`` device tensors."""
    geo = train_geometry(boxes, joints, img_w, scale_ratio, rot, flip, joint_pairs, input_shape, output_shape, want_input=True)
    heat_maps, masks = encode_heat_maps_basic(geo["joints_input"], sigma, output_shape, stride)
    return heat_maps, masks, geo["trans_inv"]


class BasicSimpleTransform(object):
    """Hot-path members of the reference's ``BasicSimpleTransform`` (commons/transforms.py:64-148)."""

    def __init__(self, joint_pairs=None, input_shape=(192, 256), output_shape=(48, 64),
                 scale=(0.7, 1.3), ratio=(-40, 40), rand_crop=True):
        self.input_shape = input_shape
        self.output_shape = output_shape
        self.joint_pairs = joint_pairs
        self.w_h_ratio = self.input_shape[0] / self.input_shape[1]
        self.scale = scale
        self.ratio = ratio
        self.rand_crop = rand_crop

    def joint_targets(self, boxes, joints, img_w=None, scale_ratio=None, rot=None, flip=None, sigma=2.0):
        """Batched ``__call__`` (reference :118-148) without ``box_crop`` and the image warp."""
        if self.joint_pairs is None:
            flip = None
        return train_targets_basic(boxes, joints, img_w, scale_ratio, rot, flip, self.joint_pairs, self.input_shape,
                                   self.output_shape, sigma)

    @staticmethod
    def get_heat_map(joints, sigma=2.0, shape=(48, 64), stride=4):
        _abi.refuse_loader_worker("BasicSimpleTransform.get_heat_map")
        arr = np.ascontiguousarray(np.asarray(joints, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("joints must be [K, 3]")
        targets, weights = encode_heat_maps_basic(arr[None], sigma, shape, stride)
        return targets[0].cpu().numpy(), weights[0].cpu().numpy()
