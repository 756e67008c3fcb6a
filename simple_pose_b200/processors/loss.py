"""Masked joints-MSE loss: drop-in for the expression every reference solver repeats
(``processors/dp_pose_hrnet_solver.py:86,106-107``)::

    loss = 0.5 * nn.MSELoss()(pred.mul(mask[[..., None, None]]), target.mul(mask[[..., None, None]]))
    loss.backward()

``JointsMSELoss()(pred, target, mask)`` returns the same 0-d float32 tensor with autograd
support. Forward and backward are ONE kernel (``sp_mse_fwd_bwd_f32``): the gradient w.r.t.
``pred`` is produced while the loss is being reduced; ``backward`` only rescales it when the
upstream gradient is not 1 (``GradScaler``), and that rescale kernel exits without touching
memory when the factor is exactly 1.
"""
import torch

from .. import _abi

def _workspace(device, stream_id):
    """Zero-initialised reduction workspace per (device, stream); the kernel restores the zero state."""
    return _abi.scratch(device, stream_id, int(_abi.lib().sp_mse_workspace_bytes()), "mse")


def mse_forward_backward(pred, target, mask, need_grad=True, grad_scale=1.0, skip_masked=False):
    """(loss 0-d float32, grad like pred or None); everything on pred's CUDA device."""
    dev = _abi.require_cuda(pred, target, mask)
    if pred.dim() < 3:
        raise ValueError("pred must be [B, K, ...]")
    b, k = int(pred.shape[0]), int(pred.shape[1])
    hw = 1
    for s in pred.shape[2:]:
        hw *= int(s)
    if tuple(target.shape) != tuple(pred.shape) or tuple(mask.shape) != (b, k):
        raise ValueError("shape mismatch: pred %s target %s mask %s" % (tuple(pred.shape), tuple(target.shape), tuple(mask.shape)))
    p = _abi.dense(pred.detach(), torch.float32)
    t = _abi.dense(target.detach(), torch.float32)
    m = _abi.dense(mask.detach(), torch.float32)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(p) if need_grad else None
    stream = _abi.stream_ptr(dev)
    ws = _workspace(dev, stream)
    flags = _abi.SP_MSE_SKIP_MASKED if skip_masked else 0
    with torch.cuda.device(dev):
        _abi.check_ws(_abi.lib().sp_mse_fwd_bwd_f32(p.data_ptr(), t.data_ptr(), m.data_ptr(), _abi.ptr(grad),
                                                    loss.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                                    b, k, hw, float(grad_scale), flags, stream), dev, stream)
    return loss, grad


class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, skip_masked):
        need = pred.requires_grad
        loss, grad = mse_forward_backward(pred, target, mask, need_grad=need, skip_masked=skip_masked)
        ctx.pred_dtype = pred.dtype
        ctx.save_for_backward(grad if need else None)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None, None, None
        dev = grad.device
        g = _abi.dense(grad_out.detach().to(dev), torch.float32)
        with torch.cuda.device(dev):
            _abi.check(_abi.lib().sp_scale_inplace_f32(grad.data_ptr(), grad.numel(), g.data_ptr(),
                                                       _abi.stream_ptr(dev)))
        out = grad if ctx.pred_dtype == torch.float32 else grad.to(ctx.pred_dtype)
        return out, None, None, None


class JointsMSELoss(torch.nn.Module):
    """``loss = JointsMSELoss()(pred [B,K,H,W], target [B,K,H,W], mask [B,K])``.

    ``skip_masked=True`` does not read pred/target of joints whose mask is 0 (saves their HBM
    traffic; differs from the reference only when those maps hold NaN/Inf, which the
    reference would propagate into the loss)."""

    def __init__(self, skip_masked=False):
        super().__init__()
        self.skip_masked = bool(skip_masked)

    def forward(self, pred, target, mask):
        return _MaskedMSE.apply(pred, target, mask, self.skip_masked)


def encode_mse_forward_backward(joints, pred, sigma=2.0, need_grad=True, want_targets=False, want_axes=False,
                                grad_scale=1.0):
    """Fused target encoding + masked MSE (+ HeatMapAcc argmaxes): one pass over ``pred``.

    joints [B,K,3] heatmap px, pred [B,K,H,W]. Returns a dict with loss (0-d), weights [B,K], and
    optionally grad, targets, pred_xy / label_xy [B,K,2]."""
    dev = _abi.require_cuda(joints, pred)
    if pred.dim() != 4:
        raise ValueError("pred must be [B, K, H, W]")
    b, k, h, w = (int(s) for s in pred.shape)
    if tuple(joints.shape) != (b, k, 3):
        raise ValueError("joints must be [B, K, 3] matching pred")
    j = _abi.dense(joints.detach(), torch.float32)
    p = _abi.dense(pred.detach(), torch.float32)
    out = {"loss": torch.empty((), dtype=torch.float32, device=dev),
           "weights": torch.empty((b, k), dtype=torch.float32, device=dev),
           "grad": torch.empty_like(p) if need_grad else None,
           "targets": torch.empty_like(p) if want_targets else None,
           "pred_xy": torch.empty((b, k, 2), dtype=torch.float32, device=dev) if want_axes else None,
           "label_xy": torch.empty((b, k, 2), dtype=torch.float32, device=dev) if want_axes else None}
    stream = _abi.stream_ptr(dev)
    ws = _workspace(dev, stream)
    if w % 4 != 0:
        # shapes the fused kernel does not take: same results from the separate kernels
        from ..commons.transforms import encode_heat_maps
        from ..metrics.pose_metrics import BasicKeyPointDecoder
        tg, wt = encode_heat_maps(j, sigma, (w, h))
        loss, grad = mse_forward_backward(p, tg, wt, need_grad=need_grad, grad_scale=grad_scale)
        out.update(loss=loss, grad=grad, weights=wt, targets=tg if want_targets else None)
        if want_axes:
            m = wt[..., None, None]
            out["pred_xy"] = BasicKeyPointDecoder.heat_map_to_axis(p * m)[0]
            out["label_xy"] = BasicKeyPointDecoder.heat_map_to_axis(tg * m)[0]
        return out
    with torch.cuda.device(dev):
        _abi.check_ws(_abi.lib().sp_encode_mse_fwd_bwd_f32(
            j.data_ptr(), p.data_ptr(), _abi.ptr(out["grad"]), _abi.ptr(out["targets"]), out["weights"].data_ptr(),
            out["loss"].data_ptr(), _abi.ptr(out["pred_xy"]), _abi.ptr(out["label_xy"]), ws.data_ptr(), ws.numel() * 8,
            b, k, h, w, float(sigma), float(grad_scale), stream), dev, stream)
    return out


class _EncodeMaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, joints, sigma, want_targets, want_axes, holder):
        need = pred.requires_grad
        out = encode_mse_forward_backward(joints, pred, sigma, need_grad=need, want_targets=want_targets,
                                          want_axes=want_axes)
        holder.update(out)
        ctx.pred_dtype = pred.dtype
        ctx.save_for_backward(out["grad"] if need else None)
        return out["loss"]

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None, None, None, None, None
        dev = grad.device
        g = _abi.dense(grad_out.detach().to(dev), torch.float32)
        with torch.cuda.device(dev):
            _abi.check(_abi.lib().sp_scale_inplace_f32(grad.data_ptr(), grad.numel(), g.data_ptr(),
                                                       _abi.stream_ptr(dev)))
        out = grad if ctx.pred_dtype == torch.float32 else grad.to(ctx.pred_dtype)
        return out, None, None, None, None, None


class EncodeJointsMSELoss(torch.nn.Module):
    """Training-step half of the path in one kernel: takes the heatmap-pixel
    joints the loader already has (``transforms.py:218``) instead of pre-rendered targets.

        crit = EncodeJointsMSELoss(sigma=2.0, with_acc=True)
        loss, acc = crit(pred, joints)          # == reference loss and HeatMapAcc()(pred*mask, target*mask)
        loss.backward()
        crit.weights                             # the [B,K] mask the reference's loader would have shipped
    """

    def __init__(self, sigma=2.0, with_acc=False, keep_targets=False, distance_thresh=0.5, norm_frac=10.):
        super().__init__()
        self.sigma, self.with_acc, self.keep_targets = float(sigma), bool(with_acc), bool(keep_targets)
        self.distance_thresh, self.norm_frac = distance_thresh, norm_frac
        self.weights = self.targets = self.acc = None

    def forward(self, pred, joints):
        holder = {}
        loss = _EncodeMaskedMSE.apply(pred, joints, self.sigma, self.keep_targets, self.with_acc, holder)
        self.weights, self.targets = holder["weights"], holder["targets"]
        if not self.with_acc:
            return loss
        from ..metrics.pose_metrics import heatmap_acc_from_axes
        self.acc = heatmap_acc_from_axes(holder["pred_xy"], holder["label_xy"], pred.shape[-2], pred.shape[-1],
                                         self.distance_thresh, self.norm_frac)
        return loss, self.acc
