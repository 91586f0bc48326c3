"""Losses -- host mirror of the reference's losses.py (losses.py:6-175) on top of librndense.so.
Same names, argument order and keyword defaults; inputs are dense `[M, C]` / `[M, K]` tensors
exactly like the reference's masked tensors.  Gradients: the reference relies on TF autodiff;
here every function is a `torch.autograd.Function` whose backward is produced by the same
kernel pass as the forward (explicit `dlogits`), so `loss.backward()` works on CUDA tensors.

`classification_loss` has the reference's committed behaviour by default (mean BCE + mean dice,
losses.py:124-139); the alternatives the reference keeps in comments (focal, balanced BCE) are
selected with the keyword-only `mode`, or globally with `CLASSIFICATION_MODE`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._tensor import as_cuda
from . import utils
from .utils import Detection  # noqa: F401  (losses.py:3)
from .pipeline import CLS_MODES

CLASSIFICATION_MODE = 'bce_dice'


def _ws(num_classes, dev):
    n = _lib.lib.rn_loss_workspace_bytes(num_classes)
    return torch.zeros((n,), dtype=torch.uint8, device=dev), n


def _two_d(labels, logits):
    labels, logits = as_cuda(labels), as_cuda(logits)
    if labels.shape != logits.shape:
        raise ValueError('labels %s and logits %s must have the same shape' % (tuple(labels.shape), tuple(logits.shape)))
    c = int(logits.shape[-1]) if logits.dim() > 0 else 1
    return labels, logits, logits.numel() // max(c, 1), c


class _Elementwise(torch.autograd.Function):
    """Unreduced focal / BCE / balanced BCE: loss [.., C] and d loss / d logits from one launch."""

    @staticmethod
    def forward(ctx, labels, logits, mode, focus, alpha, eps, w_pos, w_neg):
        labels, logits, rows, c = _two_d(labels, logits)
        loss = torch.empty_like(logits)
        grad = torch.empty_like(logits)
        _lib.call('rn_cls_loss_elementwise', rows, c, _lib.ptr(labels), _lib.ptr(logits), mode, float(focus),
                  float(alpha), float(eps), _lib.ptr(w_pos), _lib.ptr(w_neg), _lib.ptr(loss), _lib.ptr(grad),
                  _lib.stream_ptr())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g * grad, None, None, None, None, None, None


def focal_sigmoid_cross_entropy_with_logits(labels, logits, focus=2.0, alpha=0.25, eps=1e-7,
                                            name='focal_sigmoid_cross_entropy_with_logits'):
    """-alpha_t (1 - p_t)^focus log(p_t + eps), elementwise and unreduced, with the reference's
    naive `1 - p` and eps inside the log (losses.py:6-15)."""
    return _Elementwise.apply(labels, as_cuda(logits), _lib.RN_CLS_FOCAL, focus, alpha, eps, None, None)


class _FocalSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, logits, focus, alpha, eps):
        labels, logits, rows, c = _two_d(labels, logits)
        loss = torch.empty(logits.shape[:-1], dtype=torch.float32, device=logits.device)
        grad = torch.empty_like(logits)
        _lib.call('rn_focal_softmax', rows, c, _lib.ptr(labels), _lib.ptr(logits), float(focus), float(alpha),
                  float(eps), _lib.ptr(loss), _lib.ptr(grad), _lib.stream_ptr())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g.unsqueeze(-1) * grad, None, None, None


def focal_softmax_cross_entropy_with_logits(labels, logits, focus=2.0, alpha=0.25, eps=1e-7,
                                            name='focal_softmax_cross_entropy_with_logits'):
    """-sum_c a_c (1 - p_t,c)^focus z_c log(softmax(x)_c + eps), one value per row (losses.py:19-34;
    unused by the reference's classification_loss and marked "TODO: check if this is correct" there)."""
    return _FocalSoftmax.apply(labels, as_cuda(logits), focus, alpha, eps)


def sigmoid_cross_entropy_with_logits(labels, logits):
    """tf.nn.sigmoid_cross_entropy_with_logits as called at losses.py:107,124."""
    return _Elementwise.apply(labels, as_cuda(logits), _lib.RN_CLS_BCE, 2.0, 0.25, 1e-7, None, None)


def _class_stats(labels, logits=None):
    """(Z_c, I_c, P_c) float64 [C] through rn_label_stats_dense / rn_cls_sums."""
    labels = as_cuda(labels)
    c = int(labels.shape[-1])
    rows = labels.numel() // max(c, 1)
    stats = torch.zeros((_lib.RN_STATS_HEADER + 3 * c,), dtype=torch.float64, device=labels.device)
    _lib.call('rn_label_stats_dense', rows, c, _lib.ptr(labels), _lib.ptr(stats), _lib.stream_ptr())
    if logits is not None and rows > 0:
        ws, n = _ws(c, labels.device)
        seg = _lib.rows_of([rows])
        _lib.call('rn_cls_sums', C.byref(seg), c, C.byref(_lib.maps_of([logits.reshape(rows, c)])),
                  C.byref(_lib.maps_of([labels.reshape(rows, c)])), 1, _lib.RN_DTYPE_F32, _lib.ptr(stats), _lib.ptr(ws), n,
                  _lib.stream_ptr())
    h = _lib.RN_STATS_HEADER
    return stats, stats[h:h + c], stats[h + c:h + 2 * c], stats[h + 2 * c:h + 3 * c]


def balanced_sigmoid_cross_entropy_with_logits(labels, logits, axis=None,
                                               name='balanced_sigmoid_cross_entropy_with_logits'):
    """BCE weighted by n_neg/(n_pos+n_neg) on positives and n_pos/(n_pos+n_neg) on negatives,
    counts taken over `axis` (None: everything, 0: per class) (losses.py:96-110)."""
    labels, logits, rows, c = _two_d(labels, logits)
    if axis not in (None, 0):
        raise NotImplementedError('axis must be None or 0 (the reference itself notes a keepdims problem, '
                                  'losses.py:94)')
    stats, z, _, _ = _class_stats(labels)
    n_pos = z.sum().expand(c) if axis is None else z
    total = float(rows * c) if axis is None else float(rows)
    w_pos = ((total - n_pos) / total).float().contiguous()
    w_neg = (n_pos / total).float().contiguous()
    return _Elementwise.apply(labels, logits, _lib.RN_CLS_BALANCED_BCE, 2.0, 0.25, 1e-7, w_pos, w_neg)


class _ColumnSumLoss(torch.autograd.Function):
    """Losses that are a function of the per-class sums Z = sum z, I = sum z p, P = sum p."""

    @staticmethod
    def forward(ctx, labels, logits, formula, axis):
        labels, logits, rows, c = _two_d(labels, logits)
        _, z, i, p = _class_stats(labels, logits)
        out = formula(z.sum(), i.sum(), p.sum()) if axis is None else formula(z, i, p)
        # only plain tensors are kept: the tiny float64 graph of `formula` is rebuilt in backward, so a second
        # backward (retain_graph, checkpointing) works and nothing but these tensors outlives the forward
        ctx.save_for_backward(labels, logits, z, i, p)
        ctx.formula, ctx.axis = formula, axis
        return out.float()

    @staticmethod
    def backward(ctx, g):
        labels, logits, z, i, p = ctx.saved_tensors
        with torch.enable_grad():
            i_, p_ = i.detach().clone().requires_grad_(True), p.detach().clone().requires_grad_(True)
            out = ctx.formula(z.sum(), i_.sum(), p_.sum()) if ctx.axis is None else ctx.formula(z, i_, p_)
            a, b = torch.autograd.grad(out, (i_, p_), g.double().reshape(out.shape), allow_unused=True)
        c = int(logits.shape[-1])
        a = (torch.zeros_like(i_) if a is None else a).float().contiguous()
        b = (torch.zeros_like(p_) if b is None else b).float().contiguous()
        grad = torch.empty_like(logits)
        _lib.call('rn_colsum_backward', logits.numel() // max(c, 1), c, _lib.ptr(labels), _lib.ptr(logits),
                  _lib.ptr(a), _lib.ptr(b), _lib.ptr(grad), _lib.stream_ptr())
        return None, grad, None, None


def _check_axis(axis):
    if axis not in (None, 0):
        raise NotImplementedError('axis must be None or 0')


def jaccard_loss(labels, logits, smooth=1., axis=None, name='jaccard_loss'):
    """(1 - (I + s) / (U - I + s)) * s with U = sum z + sum p (losses.py:37-47)."""
    _check_axis(axis)
    return _ColumnSumLoss.apply(labels, as_cuda(logits),
                                lambda z, i, p: (1 - (i + smooth) / (z + p - i + smooth)) * smooth, axis)


def dice_loss(labels, logits, smooth=1., axis=None, name='dice_loss'):
    """1 - (2 I + s) / (sum z + sum p + s) (losses.py:50-60)."""
    _check_axis(axis)
    return _ColumnSumLoss.apply(labels, as_cuda(logits), lambda z, i, p: 1 - (2 * i + smooth) / (z + p + smooth),
                                axis)


def fixed_iou_loss(labels, logits, smooth=1., axis=0, name='fixed_iou_loss'):
    """1 - (I + s) / (sum z + sum (1 - z) p + s) (losses.py:63-73); sum (1-z) p = P - I."""
    _check_axis(axis)
    return _ColumnSumLoss.apply(labels, as_cuda(logits),
                                lambda z, i, p: 1 - (i + smooth) / (z + (p - i) + smooth), axis)


class _ClassificationLoss(torch.autograd.Function):
    """Scalar classification loss + d/d logits via the fused dense-label kernels."""

    @staticmethod
    def forward(ctx, labels, logits, fg_mask, mode):
        labels, logits, rows, c = _two_d(labels, logits)
        code = CLS_MODES[mode]
        dev = logits.device
        stats, _, _, _ = _class_stats(labels, logits if code & _lib.RN_CLS_DICE_FLAG else None)
        if fg_mask is not None:  # losses.py:120: num_fg = sum(fg_mask)
            stats[_lib.RN_STAT_NUM_FG] = as_cuda(fg_mask, torch.float64).sum()
        grad = torch.empty_like(logits)
        out = torch.zeros((4,), dtype=torch.float32, device=dev)
        if rows > 0:
            ws, n = _ws(c, dev)
            seg = _lib.rows_of([rows])
            _lib.call('rn_cls_loss_fwd_bwd', C.byref(seg), c, C.byref(_lib.maps_of([logits.reshape(rows, c)])),
                      C.byref(_lib.maps_of([labels.reshape(rows, c)])), 1, code, 2.0, 0.25, 1e-7,
                      C.byref(_lib.maps_of([grad.reshape(rows, c)])), _lib.RN_DTYPE_F32, _lib.ptr(stats), _lib.ptr(ws), n,
                      _lib.stream_ptr())
        _lib.call('rn_loss_finalize', c, code, _lib.ptr(stats), _lib.ptr(out), _lib.stream_ptr())
        ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g * grad, None, None


def classification_loss(labels, logits, fg_mask, name='classification_loss', *, mode=None):
    """mean(BCE) + mean(dice(smooth=0, axis=0)) as committed (losses.py:115-141); `mode` picks
    the commented-out variants: 'focal' (sum / max(num_fg, 1)), 'balanced_bce', ... ."""
    mode = CLASSIFICATION_MODE if mode is None else mode
    if mode not in CLS_MODES:
        raise ValueError('mode must be one of %s' % sorted(CLS_MODES))
    return _ClassificationLoss.apply(labels, as_cuda(logits), fg_mask, mode)


class _RegressionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, logits, fg_mask):
        labels, logits = as_cuda(labels), as_cuda(logits)
        if labels.shape != logits.shape or logits.dim() != 2:
            raise ValueError('labels and logits must both be [M, K]')
        rows, width = int(logits.shape[0]), int(logits.shape[1])
        fg = as_cuda(fg_mask, torch.uint8).reshape(-1)
        if fg.shape[0] != rows:
            raise ValueError('fg_mask must be [M]')
        dev = logits.device
        stats = torch.zeros((_lib.RN_STATS_HEADER + 3,), dtype=torch.float64, device=dev)
        stats[_lib.RN_STAT_NUM_FG] = fg.sum(dtype=torch.float64)
        out = torch.zeros((1,), dtype=torch.float32, device=dev)
        grad = torch.zeros_like(logits)
        ws, n = _ws(1, dev)
        _lib.call('rn_huber_rows', rows, width, _lib.ptr(logits), _lib.ptr(labels), _lib.ptr(fg), _lib.ptr(grad),
                  _lib.ptr(stats), _lib.ptr(out), _lib.ptr(ws), n, _lib.stream_ptr())
        ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, g * grad, None


def regression_loss(labels, logits, fg_mask, name='regression_loss'):
    """tf.losses.huber_loss(labels, predictions=logits, weights=fg_mask[:, None],
    reduction=SUM_BY_NONZERO_WEIGHTS): Huber(delta=1) summed over foreground anchors, divided
    by the number of non-zero broadcast weights, 0 when there is none (losses.py:144-152)."""
    return _RegressionLoss.apply(labels, as_cuda(logits), fg_mask)


def loss(labels: Detection, logits: Detection, name='loss', *, mode=None):
    """(class_loss, regr_loss) of masked `Detection`s (losses.py:155-175).  The reference's two
    tf.summary.histogram calls are logging and have no counterpart."""
    fg_mask = utils.classmap_decode(labels.classification.prob).fg_mask
    class_loss = classification_loss(labels=labels.classification.prob, logits=logits.classification.unscaled,
                                     fg_mask=fg_mask, mode=mode)
    regr_loss = regression_loss(labels=labels.regression, logits=logits.regression, fg_mask=fg_mask)
    return class_loss, regr_loss
