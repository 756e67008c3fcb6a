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


class RefineSimpleTransform(object):
    """Only the hot-path member of the reference class is mirrored; augmentation
    (``__call__``) stays with the reference's CPU loader."""

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
        """joints [K,3] ndarray -> (targets [K,H,W] float32 ndarray, weights [K] float32 ndarray)."""
        arr = np.ascontiguousarray(np.asarray(joints, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("joints must be [K, 3]")
        targets, weights = encode_heat_maps(arr[None], sigma, shape)
        return targets[0].cpu().numpy(), weights[0].cpu().numpy()


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


def encode_heat_maps_basic(joints, sigma=2.0, shape=(48, 64), stride=4):
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
    targets = torch.empty((b, k, height, width), dtype=torch.float32, device=dev)
    weights = torch.empty((b, k), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(_abi.lib().sp_encode_basic_f32(j.data_ptr(), table.data_ptr(), targets.data_ptr(), weights.data_ptr(),
                                                  b, k, height, width, float(sigma), int(stride), int(table.shape[0]),
                                                  _abi.stream_ptr(dev)))
    return targets, weights


class BasicSimpleTransform(object):
    """Hot-path member of the reference's ``BasicSimpleTransform`` (commons/transforms.py:64-116)."""

    @staticmethod
    def get_heat_map(joints, sigma=2.0, shape=(48, 64), stride=4):
        arr = np.ascontiguousarray(np.asarray(joints, dtype=np.float32))
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("joints must be [K, 3]")
        targets, weights = encode_heat_maps_basic(arr[None], sigma, shape, stride)
        return targets[0].cpu().numpy(), weights[0].cpu().numpy()
