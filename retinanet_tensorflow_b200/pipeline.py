"""Batched, fused dense-head path on one B200: assign+encode -> loss fwd+bwd -> decode/select/NMS.

`DenseHeadPath` owns every buffer of the path (compact labels, statistics vector, gradient
buffers, candidate / detection lists, workspaces) and issues the kernels of librndense.so on
the current CUDA stream; nothing synchronises with the host.  It is the fast equivalent of the
reference's three call sites: `dataset.build_labels` per image on tf.data threads
(dataset.py:167), `utils.process_labels_and_logits` + `losses.loss` in the training graph
(train.py:218-220) and `utils.boxes_decode` + `nms_classwise` (train.py:68-75).

Per-anchor tensors use the merged level-major layout: the concatenation over levels P3..P7
of the reference's per-level `[B, H, W, A, K]` tensors (utils.merge_outputs order,
utils.py:151-153).  A list of per-level tensors is accepted wherever a merged tensor is.

Multi-GPU: one process per GPU, the batch sharded by image.  The only exchange is an
all-reduce(sum) of slices of the float64 statistics vector (normalisers before the gradient
pass, loss sums after it) through `torch.distributed`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, dist
from ._pyramid import level_rows, level_shapes, pyramid_for, split_merged
from .levels import build_levels

CLS_MODES = {
    'focal': _lib.RN_CLS_FOCAL,
    'bce': _lib.RN_CLS_BCE,
    'balanced_bce': _lib.RN_CLS_BALANCED_BCE,
    'bce_dice': _lib.RN_CLS_BCE | _lib.RN_CLS_DICE_FLAG,                    # the reference as committed
    'balanced_bce_dice': _lib.RN_CLS_BALANCED_BCE | _lib.RN_CLS_DICE_FLAG,  # README.md:27
    'focal_dice': _lib.RN_CLS_FOCAL | _lib.RN_CLS_DICE_FLAG,
    'dice': _lib.RN_CLS_NONE | _lib.RN_CLS_DICE_FLAG,
}

_FLAG_MESSAGES = {
    _lib.RN_FLAG_BOX_ORDER: 'assert_less_equal(box[:2], box[2:]) failed (utils.py:66-67)',
    _lib.RN_FLAG_IOU_RANGE: 'IoU outside [0, 1] or NaN (utils.py:94)',
    _lib.RN_FLAG_CAND_OVERFLOW: 'more candidates than `capacity`',
    _lib.RN_FLAG_DET_OVERFLOW: 'more detections than `out_capacity`',
    _lib.RN_FLAG_EMPTY_GT: 'an image has no ground-truth box (data_loaders/inferred.py:30)',
    _lib.RN_FLAG_EXCHANGE_TIMEOUT: 'a peer rank did not post its statistics in time: the statistics (and every '
                                   'loss / gradient that depends on them) are NaN',
}


def raise_for_status(word, ignore=0):
    """Turn the device status word into the exception the reference would raise."""
    word = int(word) & ~ignore
    if word & (_lib.RN_FLAG_BOX_ORDER | _lib.RN_FLAG_IOU_RANGE | _lib.RN_FLAG_EMPTY_GT):
        raise AssertionError('; '.join(m for f, m in _FLAG_MESSAGES.items() if word & f))
    if word:
        raise ValueError('; '.join(m for f, m in _FLAG_MESSAGES.items() if word & f))


class DenseHeadPath(object):
    def __init__(self, image_size, num_classes, batch, max_gt, levels=None, *, cls_mode='focal',
                 anchor_size_mode='tf1_int_trunc', neg_iou=0.4, pos_iou=0.5, focus=2.0, alpha=0.25, eps=1e-7,
                 score_thr=0.5, nms_iou=0.5, max_output_per_class=1000, topk_per_level=0, capacity=4096,
                 out_capacity=None, keep_iou=False, device=None, process_group=None, exchange='auto',
                 generic_kernels=False, assign_patch=(0, 0), one_launch_tail=True):
        self.device = _lib.require_cuda(device)
        # sub-patches (rows, columns) one CTA of the assign kernel handles; (0, 0) = the library's choice.  The
        # results do not depend on it (tests/test_gpu_assign.py runs several decompositions).
        self.assign_patch = (int(assign_patch[0]), int(assign_patch[1]))
        # one_launch_tail=False keeps candidate finalize and NMS as separate launches (the parity tests compare both)
        self.one_launch_tail = bool(one_launch_tail)
        self.levels = build_levels() if levels is None else levels
        self.names = list(self.levels)
        self.image_size = (int(image_size[0]), int(image_size[1]))
        self.num_classes = int(num_classes)
        self.batch = int(batch)
        self.max_gt = int(max_gt)
        if cls_mode not in CLS_MODES:
            raise ValueError('cls_mode must be one of %s' % sorted(CLS_MODES))
        self.cls_mode = cls_mode
        # generic_kernels=True: this path object takes the generic loss kernel (libm, IEEE division) even where the
        # streaming kernel applies -- a per-object choice (RN_CLS_GENERIC_FLAG per call), used by the parity tests
        self.mode = CLS_MODES[cls_mode] | (_lib.RN_CLS_GENERIC_FLAG if generic_kernels else 0)
        self.neg_iou, self.pos_iou = float(neg_iou), float(pos_iou)
        self.focus, self.alpha, self.eps = float(focus), float(alpha), float(eps)
        self.score_thr, self.nms_iou = float(score_thr), float(nms_iou)
        self.max_output_per_class = int(max_output_per_class)
        self.topk_per_level = int(topk_per_level or 0)
        self.capacity = int(capacity)
        self.out_capacity = int(out_capacity if out_capacity is not None else capacity)
        self.group = process_group
        self.exchange = exchange   # 'nccl' | 'peer' (own kernel over NVLink peer memory) | 'auto'
        self._peer = None
        self.stage_hook = None  # optional callable(stage_name, thunk) -> thunk(); bench.py times stages with it
        self._side = None        # side stream for the detection tail (step(overlap_detect=True))
        self._side_done = None
        self.pyr = pyramid_for(self.image_size, self.levels, anchor_size_mode, self.names)
        self.anchors = int(self.pyr.anchors_per_image)
        self.rows_total = self.batch * self.anchors
        self.rows = _lib.Rows()
        _lib.call('rn_rows_from_pyramid', C.byref(self.pyr), self.batch, C.byref(self.rows))

        dev, B, R, Cn = self.device, self.batch, self.rows_total, self.num_classes
        i32, f32 = torch.int32, torch.float32
        self._keep_iou = bool(keep_iou)
        # label buffers (K1 output, K2 input); a second set appears with the first prefetch_assign()
        self._sets = [None, None]
        self._cur = 0
        self._pending = None          # (set index, stats index) of an assignment launched ahead of its step
        self._assign_stream = None
        self._assign_done = None
        self._loss_done = [None, None]   # events after the loss kernels of the last two steps (newest first)
        # Three statistics vectors used in rotation: step i accumulates into ring[i % 3]; its loss kernel leaves
        # ring[(i + 2) % 3] zeroed (last used by step i - 1, whose side-stream tail has been waited for by then; the
        # vector of step i + 1 may already be filling: prefetch_assign), so a training loop needs neither a memset
        # nor a snapshot copy per step.
        self._stats_ring = torch.zeros((3, _lib.RN_STATS_HEADER + 3 * Cn), dtype=torch.float64, device=dev)
        self._stats_clean = [True, True, True]
        self._stats_i = 0
        self._static_stats = False   # graph capture: one fixed vector, zeroed by an op of its own
        self.stats = self._stats_ring[0]
        self.loss_out = torch.zeros((4,), dtype=f32, device=dev)
        self.status = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._loss_ws_bytes = _lib.lib.rn_loss_workspace_bytes(Cn)
        self._loss_ws = torch.zeros((self._loss_ws_bytes,), dtype=torch.uint8, device=dev)
        self._huber_ws = torch.zeros((_lib.lib.rn_loss_workspace_bytes(1),), dtype=torch.uint8, device=dev)
        self.dlogits = None
        self.dreg = None
        cap, ocap = self.capacity, self.out_capacity
        self.cand_boxes = torch.empty((B, cap, 4), dtype=f32, device=dev)
        self.cand_scores = torch.empty((B, cap), dtype=f32, device=dev)
        self.cand_classes = torch.empty((B, cap), dtype=i32, device=dev)
        self.cand_anchor = torch.empty((B, cap), dtype=i32, device=dev)
        self.cand_count = torch.zeros((B,), dtype=i32, device=dev)
        self.det_boxes = torch.empty((B, ocap, 4), dtype=f32, device=dev)
        self.det_scores = torch.empty((B, ocap), dtype=f32, device=dev)
        self.det_classes = torch.empty((B, ocap), dtype=i32, device=dev)
        self.det_index = torch.empty((B, ocap), dtype=i32, device=dev)
        self.det_count = torch.zeros((B,), dtype=i32, device=dev)
        self._sel_ws_bytes = _lib.lib.rn_select_workspace_bytes(B, cap)
        self._sel_ws = torch.empty((self._sel_ws_bytes,), dtype=torch.uint8, device=dev)
        self._nms_ws_bytes = _lib.lib.rn_nms_workspace_bytes(B, Cn, cap)
        self._nms_ws = torch.empty((self._nms_ws_bytes,), dtype=torch.uint8, device=dev)

        self._use_set(0)

    # ---- layout helpers ----------------------------------------------------------------------
    def _use_set(self, k):
        """Make label-buffer set k the current one (allocating it on first use)."""
        if self._sets[k] is None:
            R, dev = self.rows_total, self.device
            t = dict(matched=torch.empty((R,), dtype=torch.int32, device=dev),
                     label=torch.empty((R,), dtype=torch.int32, device=dev),
                     deltas=torch.empty((R, 4), dtype=torch.float32, device=dev),
                     iou_max=torch.empty((R,), dtype=torch.float32, device=dev) if self._keep_iou else None)
            t['maps'] = (self._merged_maps(t['matched'], 4), self._merged_maps(t['label'], 4),
                         self._merged_maps(t['deltas'], 16), self._merged_maps(t['iou_max'], 4))
            self._sets[k] = t
        t = self._sets[k]
        self._cur = k
        self.matched, self.label, self.deltas, self.iou_max = t['matched'], t['label'], t['deltas'], t['iou_max']
        self._m_matched, self._m_label, self._m_deltas, self._m_iou = t['maps']

    def _merged_maps(self, tensor, bytes_per_anchor):
        m = _lib.Maps()
        if tensor is not None:
            _lib.call('rn_maps_from_merged', C.byref(self.pyr), self.batch, _lib.ptr(tensor), bytes_per_anchor,
                      C.byref(m))
        return m

    def _maps(self, x, trailing, dtypes=(torch.float32,)):
        """rn_maps_t for a merged tensor [R, trailing] or a list/dict of per-level tensors."""
        if isinstance(x, dict):
            x = [x[n] for n in self.names]
        if isinstance(x, (list, tuple)):
            shapes = level_shapes(self.pyr)
            if len(x) != len(shapes):
                raise ValueError('expected %d levels' % len(shapes))
            for t, (h, w, a) in zip(x, shapes):
                if tuple(t.shape) != (self.batch, h, w, a, trailing) or t.dtype not in dtypes or t.dtype != x[0].dtype:
                    raise ValueError('level tensor must be %s [%d,%d,%d,%d,%d], got %s %s'
                                     % (dtypes, self.batch, h, w, a, trailing, tuple(t.shape), t.dtype))
            ts = [t.contiguous() for t in x]
            return _lib.maps_of(ts), ts
        if tuple(x.shape) != (self.rows_total, trailing) or x.dtype not in dtypes:
            raise ValueError('merged tensor must be %s [%d,%d], got %s %s'
                             % (dtypes, self.rows_total, trailing, tuple(x.shape), x.dtype))
        x = x.contiguous()
        return self._merged_maps(x, x.element_size() * trailing), x

    _LOGIT_DTYPES = (torch.float32, torch.bfloat16)

    def _grad_buffers(self, logits_t, want_grad):
        """(dlogits maps, dreg maps, logits dtype code); dlogits follows the logits' dtype."""
        first = logits_t[0] if isinstance(logits_t, (list, tuple)) else logits_t
        code = _lib.RN_DTYPE_BF16 if first.dtype == torch.bfloat16 else _lib.RN_DTYPE_F32
        if not want_grad:
            return _lib.Maps(), _lib.Maps(), code
        Cn = self.num_classes
        if self.dlogits is None or self.dlogits.dtype != first.dtype:
            self.dlogits = torch.empty((self.rows_total, Cn), dtype=first.dtype, device=self.device)
        if self.dreg is None:
            self.dreg = torch.empty((self.rows_total, 4), dtype=torch.float32, device=self.device)
        return (self._merged_maps(self.dlogits, self.dlogits.element_size() * Cn), self._merged_maps(self.dreg, 16),
                code)

    def per_level(self, merged, trailing=()):
        """Reference-shaped views `[B, H, W, A, ...]` keyed 'P3'.. of a merged buffer."""
        return dict(zip(self.names, split_merged(self.pyr, self.batch, merged, trailing)))

    @staticmethod
    def per_level_static(pyr, batch, merged, width):
        """Per-level `[B, H, W, A, width]` views of a merged buffer for an explicit pyramid/batch."""
        return split_merged(pyr, batch, merged, (width,))

    def _multi_rank(self):
        d = torch.distributed
        return self.group is not False and d.is_available() and d.is_initialized() and d.get_world_size(
            self.group if self.group is not None else None) > 1

    def _allreduce(self, phase, stats=None):
        """Sum one slice of the stats vector over the ranks ('counts' | 'dice_sums' | 'loss_sums').
        process_group=False disables.  exchange='peer' uses rn_stats_exchange (this library's kernel
        over NVLink peer memory), 'nccl' torch.distributed.all_reduce, 'auto' tries peer first."""
        if not self._multi_rank():
            return
        if self.exchange in ('peer', 'auto') and self._peer is None:
            try:
                self._peer = dist.PeerExchange(self.num_classes, self.device, self.group)
            except Exception as err:  # symmetric memory unavailable: fall back to the collective library
                if self.exchange == 'peer':
                    raise
                import warnings
                warnings.warn('peer-memory exchange unavailable (%s); using NCCL all-reduce' % (err,))
                self.exchange = 'nccl'
        stats = self.stats if stats is None else stats
        if self._peer is not None and self.exchange != 'nccl':
            self._peer.all_reduce(stats, phase, self.status)
        else:
            dist.all_reduce_sum(stats[dist.slices(self.num_classes)[phase]], self.group)

    def _stage(self, name, fn):
        return fn() if self.stage_hook is None else self.stage_hook(name, fn)

    # ---- stage 1: assignment + encoding (dataset.build_labels) ------------------------------
    def _second_half(self, maps, bytes_per_anchor, images):
        """The per-level pointers of `maps` advanced past the first `images` images of every level."""
        out = _lib.Maps()
        for l, (h, w, a) in enumerate(level_shapes(self.pyr)):
            out.p[l] = maps.p[l] + images * h * w * a * bytes_per_anchor if maps.p[l] else None
        return out

    def _rotate_stats(self):
        """self.stats <- the next vector of the ring, zero (by an earlier loss kernel, else by an op here)."""
        if not self._static_stats:
            self._stats_i = (self._stats_i + 1) % 3
            self.stats = self._stats_ring[self._stats_i]
        if self._static_stats or not self._stats_clean[self._stats_i]:
            self.stats.zero_()
        self._stats_clean[self._stats_i] = False

    def _mark_loss_done(self):
        """Event after this step's loss kernel: a prefetched assignment of step i + 2 overwrites the label set and
        reuses the statistics vector this kernel has just read / zeroed."""
        if self._assign_stream is not None and not torch.cuda.is_current_stream_capturing():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._loss_done = [ev, self._loss_done[0]]

    def _stats_next_ptr(self):
        """The vector the loss kernel may zero for the next step (None in static mode)."""
        if self._static_stats:
            return None
        nxt = (self._stats_i + 2) % 3
        self._stats_clean[nxt] = True
        return _lib.ptr(self._stats_ring[nxt])

    def assign(self, gt_boxes, gt_classes, gt_counts, with_flip=False):
        """gt_boxes f32 [B, max_gt, 4], gt_classes i32 [B, max_gt], gt_counts i32 [B] (device).
        Fills self.matched / label / deltas and the label statistics.

        with_flip=True builds the reference's `[orig, flipped]` sample (dataset.py:182-204) on the
        device: the ground truth holds batch/2 images, they are assigned into the first half of every
        level and their horizontally flipped label maps (augmentation.py:5-22) into the second half,
        so that images b and b + batch/2 are a flip pair (the caller flips the pixels)."""
        B = self.batch
        if with_flip:
            if B % 2:
                raise ValueError('with_flip needs an even batch (it holds [orig, flipped])')
            B //= 2
        if tuple(gt_boxes.shape) != (B, self.max_gt, 4) or gt_boxes.dtype != torch.float32:
            raise ValueError('gt_boxes must be float32 [%d, %d, 4]' % (B, self.max_gt))
        if tuple(gt_classes.shape) != (B, self.max_gt) or gt_classes.dtype != torch.int32:
            raise ValueError('gt_classes must be int32 [%d, %d]' % (B, self.max_gt))
        if tuple(gt_counts.shape) != (B,) or gt_counts.dtype != torch.int32:
            raise ValueError('gt_counts must be int32 [%d]' % B)
        if self._pending is not None:   # an assignment was launched ahead: this call supersedes it
            torch.cuda.current_stream().wait_event(self._assign_done)
            self._pending = None
        self._rotate_stats()
        gt_boxes, gt_classes, gt_counts = gt_boxes.contiguous(), gt_classes.contiguous(), gt_counts.contiguous()
        self._stage('assign', lambda: _lib.call(
            'rn_assign_encode_patch', C.byref(self.pyr), B, _lib.ptr(gt_boxes), _lib.ptr(gt_classes),
            _lib.ptr(gt_counts), self.max_gt, self.num_classes, self.neg_iou, self.pos_iou, C.byref(self._m_matched),
            C.byref(self._m_label), C.byref(self._m_deltas), C.byref(self._m_iou), _lib.ptr(self.stats),
            _lib.ptr(self.status), self.assign_patch[0], self.assign_patch[1], _lib.stream_ptr()))
        if with_flip:
            st = _lib.stream_ptr()
            none = _lib.Maps()
            _lib.call('rn_flip_labels', C.byref(self.pyr), B, C.byref(self._m_label), C.byref(self._m_deltas),
                      C.byref(self._second_half(self._m_label, 4, B)),
                      C.byref(self._second_half(self._m_deltas, 16, B)), st)
            for m in (self._m_matched, self._m_iou):   # 4-byte payloads: the label slot of the same kernel moves them
                if m.p[0]:
                    _lib.call('rn_flip_labels', C.byref(self.pyr), B, C.byref(m), C.byref(none),
                              C.byref(self._second_half(m, 4, B)), C.byref(none), st)
            self.stats[:_lib.RN_STATS_HEADER + self.num_classes] *= 2.0   # a flip keeps every count
        self._allreduce('counts')
        return self.label, self.deltas

    def prefetch_assign(self, gt_boxes, gt_classes, gt_counts, with_flip=False):
        """Launch the NEXT step's assignment now, on its own stream, into the other label-buffer set: K1 is bound
        by float32 ALU work and the loss kernel by HBM, so the two share the SMs well and the assignment (and, with
        several GPUs, its count exchange) leaves the step's critical path.  Call it right after `step()` /
        `loss()` of the current step; the next `step()` (whose ground-truth arguments are then ignored) picks the
        result up.  The ground-truth tensors must stay unchanged until that step has been issued."""
        main = torch.cuda.current_stream()
        if self._static_stats:
            raise RuntimeError('prefetch_assign is not available on a captured (static) path')
        if self._pending is not None:
            main.wait_event(self._assign_done)
            self._pending = None
        if self._assign_stream is None:
            self._assign_stream = torch.cuda.Stream(device=self.device)
            self._assign_stream.wait_stream(main)        # first use: no per-step events yet
        elif self._loss_done[1] is not None:
            self._assign_stream.wait_event(self._loss_done[1])   # loss kernel of the step before the current one
        else:
            self._assign_stream.wait_stream(main)
        cur_set, cur_stats, cur_i = self._cur, self.stats, self._stats_i
        try:
            with torch.cuda.stream(self._assign_stream):
                self._use_set(cur_set ^ 1)
                self.assign(gt_boxes, gt_classes, gt_counts, with_flip=with_flip)   # rotates the statistics vector
                self._assign_done = torch.cuda.Event()
                self._assign_done.record(self._assign_stream)
            self._pending = (cur_set ^ 1, self._stats_i)
        finally:   # the current step's buffers stay current until step() switches
            self._use_set(cur_set)
            self.stats, self._stats_i = cur_stats, cur_i
        return self._pending

    def _take_prefetched(self):
        """Switch to the label set / statistics vector of a prefetched assignment (step() calls this)."""
        set_k, stats_i = self._pending
        self._pending = None
        torch.cuda.current_stream().wait_event(self._assign_done)
        self._use_set(set_k)
        self._stats_i = stats_i
        self.stats = self._stats_ring[stats_i]

    # ---- stage 2: loss forward + backward (utils.process_labels_and_logits + losses.loss) ---
    def loss(self, logits, regression, want_grad=True):
        """Returns self.loss_out (device f32 [4]: class loss, regression loss, num_fg, M); the
        gradients w.r.t. the full logits / regression maps land in self.dlogits / self.dreg."""
        self.join()
        m_logits, logits_t = self._maps(logits, self.num_classes, self._LOGIT_DTYPES)
        m_reg, _ = self._maps(regression, 4)
        st = _lib.stream_ptr()
        Cn = self.num_classes
        m_dlogits, m_dreg, dt = self._grad_buffers(logits_t, want_grad)
        if self.mode & _lib.RN_CLS_DICE_FLAG:
            self._stage('cls_sums', lambda: _lib.call(
                'rn_cls_sums', C.byref(self.rows), Cn, C.byref(m_logits), C.byref(self._m_label), 0, dt,
                _lib.ptr(self.stats), _lib.ptr(self._loss_ws), self._loss_ws_bytes, st))
            self._allreduce('dice_sums')
        single = not self._multi_rank()
        self._stage('cls_loss', lambda: _lib.call(
            'rn_head_loss', C.byref(self.rows), Cn, C.byref(m_logits), C.byref(self._m_label), self.mode, self.focus,
            self.alpha, self.eps, C.byref(m_dlogits), dt, C.byref(m_reg), C.byref(self._m_deltas), C.byref(m_dreg),
            _lib.ptr(self.stats), self._stats_next_ptr(), _lib.ptr(self._loss_ws), self._loss_ws_bytes,
            _lib.ptr(self.loss_out) if single else None, 0.0, 0, None, _lib.ptr(self.status), None, 0, st))
        self._mark_loss_done()
        if not single:
            self._finish_loss(self.stats)
        return self.loss_out

    # ---- stages 2+3 fused: one pass over the logits for the loss AND the candidate selection --
    def fused_select_supported(self):
        return bool(_lib.lib.rn_fused_select_supported(self.num_classes, self.mode, self.focus, C.byref(self.rows))) \
            and 0.0 < self.score_thr < 1.0

    def _can_defer_loss(self):
        """The loss-sum exchange + finalize may leave the critical path (side stream) unless the exchange is an
        NCCL collective, whose cross-stream ordering is torch's business, not ours."""
        return not self._multi_rank() or (self._peer is not None and self.exchange != 'nccl')

    def _finish_loss(self, stats):
        self._allreduce('loss_sums', stats)
        _lib.call('rn_loss_finalize', self.num_classes, self.mode, _lib.ptr(stats), _lib.ptr(self.loss_out),
                  _lib.stream_ptr())

    def loss_and_select(self, logits, regression, want_grad=True, finalize=True, defer_loss=False):
        """`loss()` and `select()` on the same logits with the [rows, C] tensor read ONCE: the
        focal kernel also pre-selects the anchors whose best class probability exceeds score_thr
        (training step that also produces detections, train.py:218-220 + 68-75).
        finalize=False leaves the candidates unordered in the workspace (call select_finalize)."""
        if not self.fused_select_supported():
            self.loss(logits, regression, want_grad)
            self.select(logits, regression)
            return self.loss_out, self.cand_count
        self.join()
        m_logits, logits_t = self._maps(logits, self.num_classes, self._LOGIT_DTYPES)
        m_reg, _ = self._maps(regression, 4)
        st = _lib.stream_ptr()
        Cn = self.num_classes
        m_dlogits, m_dreg, dt = self._grad_buffers(logits_t, want_grad)
        if self.mode & _lib.RN_CLS_DICE_FLAG:
            self._stage('cls_sums', lambda: _lib.call(
                'rn_cls_sums', C.byref(self.rows), Cn, C.byref(m_logits), C.byref(self._m_label), 0, dt,
                _lib.ptr(self.stats), _lib.ptr(self._loss_ws), self._loss_ws_bytes, st))
            self._allreduce('dice_sums')
        # single GPU and nothing deferred: the loss kernel's last CTA also writes the final scalars
        in_kernel = not self._multi_rank() and not defer_loss
        self._stage('cls_loss', lambda: _lib.call(
            'rn_head_loss', C.byref(self.rows), Cn, C.byref(m_logits), C.byref(self._m_label), self.mode, self.focus,
            self.alpha, self.eps, C.byref(m_dlogits), dt, C.byref(m_reg), C.byref(self._m_deltas), C.byref(m_dreg),
            _lib.ptr(self.stats), self._stats_next_ptr(), _lib.ptr(self._loss_ws), self._loss_ws_bytes,
            _lib.ptr(self.loss_out) if in_kernel else None, self.score_thr, self.capacity, _lib.ptr(self.cand_count),
            _lib.ptr(self.status), _lib.ptr(self._sel_ws), self._sel_ws_bytes, st))
        self._mark_loss_done()
        if not in_kernel and not defer_loss:
            self._finish_loss(self.stats)
        if finalize:
            self.select_finalize(regression)
        return self.loss_out, self.cand_count

    def select_finalize(self, regression):
        """Order / de-duplicate / decode the candidates the fused pass left in the workspace."""
        m_reg, _ = self._maps(regression, 4)
        self._stage('select', lambda: _lib.call(
            'rn_select_finalize', C.byref(self.pyr), self.batch, C.byref(m_reg), 1, self.topk_per_level,
            self.capacity, 1, _lib.ptr(self.cand_boxes), _lib.ptr(self.cand_scores), _lib.ptr(self.cand_classes),
            _lib.ptr(self.cand_anchor), _lib.ptr(self.cand_count), _lib.ptr(self._sel_ws), self._sel_ws_bytes,
            _lib.stream_ptr()))
        return self.cand_count

    def tail_fused_supported(self):
        """The detection tail (candidate finalize + class-wise NMS) runs as ONE launch when the candidate lists fit
        the kernel's shared memory (rn_detect_tail: capacity <= 8192)."""
        return self.one_launch_tail and bool(_lib.lib.rn_detect_tail_supported(self.num_classes, self.capacity))

    def detect_tail(self, regression, dedup=True, regs_are_deltas=True):
        """select_finalize() + nms() in one launch; same results (tests/test_gpu_decode.py)."""
        if not self.tail_fused_supported():
            self.select_finalize(regression)
            return self.nms()
        m_reg, _ = self._maps(regression, 4)
        self._stage('tail', lambda: _lib.call(
            'rn_detect_tail', C.byref(self.pyr), self.batch, self.num_classes, C.byref(m_reg), int(regs_are_deltas),
            self.capacity, int(dedup), self.topk_per_level, _lib.ptr(self.cand_boxes), _lib.ptr(self.cand_scores), _lib.ptr(self.cand_classes),
            _lib.ptr(self.cand_anchor), _lib.ptr(self.cand_count), self.nms_iou, self.max_output_per_class,
            self.out_capacity, _lib.ptr(self.det_boxes), _lib.ptr(self.det_scores), _lib.ptr(self.det_classes),
            _lib.ptr(self.det_index), _lib.ptr(self.det_count), _lib.ptr(self.status), _lib.ptr(self._sel_ws),
            self._sel_ws_bytes, _lib.stream_ptr()))
        return self.det_count

    # ---- stage 3: decode + select + class-wise NMS (utils.boxes_decode + nms_classwise) -----
    def select(self, scores, regression, scores_are_logits=True, regs_are_deltas=True, finalize=True):
        """Candidates of utils.boxes_decode over all levels.  finalize=False runs only the select pass and leaves the
        candidates unordered in the workspace (finish with detect_tail(dedup=False))."""
        self.join()
        m_scores, scores_t = self._maps(scores, self.num_classes, self._LOGIT_DTYPES)
        first = scores_t[0] if isinstance(scores_t, (list, tuple)) else scores_t
        fmt = (_lib.RN_SCORES_LOGITS if scores_are_logits else 0) | \
            (_lib.RN_SCORES_BF16 if first.dtype == torch.bfloat16 else 0)
        m_reg, _ = self._maps(regression, 4)
        self._stage('select', lambda: _lib.call(
            'rn_decode_select', C.byref(self.pyr), self.batch, self.num_classes, C.byref(m_scores),
            fmt, C.byref(m_reg), int(regs_are_deltas), self.score_thr, self.topk_per_level if finalize else -1,
            self.capacity, _lib.ptr(self.cand_boxes), _lib.ptr(self.cand_scores), _lib.ptr(self.cand_classes),
            _lib.ptr(self.cand_anchor), _lib.ptr(self.cand_count), _lib.ptr(self.status), _lib.ptr(self._sel_ws),
            self._sel_ws_bytes, _lib.stream_ptr()))
        return self.cand_count

    def nms(self):
        self.join()
        self._stage('nms', lambda: _lib.call(
            'rn_nms_classwise', self.batch, self.num_classes, self.capacity, _lib.ptr(self.cand_boxes),
            _lib.ptr(self.cand_scores), _lib.ptr(self.cand_classes), _lib.ptr(self.cand_count), self.nms_iou,
            self.max_output_per_class, self.out_capacity, _lib.ptr(self.det_boxes), _lib.ptr(self.det_scores),
            _lib.ptr(self.det_classes), _lib.ptr(self.det_index), _lib.ptr(self.det_count), _lib.ptr(self.status),
            _lib.ptr(self._nms_ws), self._nms_ws_bytes, _lib.stream_ptr()))
        return self.det_count

    def detect(self, scores, regression, scores_are_logits=True, regs_are_deltas=True):
        """Inference post-processing: select pass over the scores + ONE launch for ordering, optional per-level top-k,
        box decoding and class-wise NMS (separate launches where the one-launch tail does not apply)."""
        if self.tail_fused_supported():
            self.select(scores, regression, scores_are_logits, regs_are_deltas, finalize=False)
            self.detect_tail(regression, dedup=False, regs_are_deltas=regs_are_deltas)
        else:
            self.select(scores, regression, scores_are_logits, regs_are_deltas)
            self.nms()
        return self.det_boxes, self.det_scores, self.det_classes, self.det_count

    # ---- one whole step of the hot path -----------------------------------------------------
    def step(self, gt_boxes, gt_classes, gt_counts, logits, regression, fused=True, overlap_detect=False,
             with_flip=False):
        """assign -> loss fwd+bwd -> decode/select/NMS.  With `fused` (default) and a supported
        configuration the loss kernel also does the candidate selection (logits read once).

        overlap_detect=True puts the latency-bound tail (candidate finalize + NMS, one CTA per
        image) on a side stream so that it overlaps the next step's assignment and loss kernels;
        events keep the select workspace from being overwritten early.  The loss-sum exchange and
        the final normalisation (`loss_out`) ride on the side stream too (unless the exchange is an
        NCCL collective).  Detections and `loss_out` are then valid after `join()` (or a device
        synchronize); gradients are on the current stream as always.

        with_flip=True: the ground truth holds batch/2 images and the batch is the reference's
        `[orig, flipped]` sample (see `assign`)."""
        if self._pending is not None:
            self._take_prefetched()
        else:
            self.assign(gt_boxes, gt_classes, gt_counts, with_flip=with_flip)
        if not (fused and self.fused_select_supported()):
            self.join()
            self.loss(logits, regression)
            self.detect(logits, regression)   # select pass + one-launch tail (or the separate launches)
            return self.loss_out, self.det_count
        main = torch.cuda.current_stream()
        if self._side_done is not None:
            main.wait_event(self._side_done)   # previous tail has consumed the select workspace
            self._side_done = None
        if not overlap_detect:
            self.loss_and_select(logits, regression, finalize=False)
            self.detect_tail(regression)
            return self.loss_out, self.det_count
        defer = self._multi_rank() and self._can_defer_loss() and not self._static_stats
        self.loss_and_select(logits, regression, finalize=False, defer_loss=defer)
        stats_now = self.stats
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)  # small CTAs first
        if not torch.cuda.is_current_stream_capturing():
            # the tail reads the caller's regression tensor on the side stream: keep the caching allocator from
            # handing its memory to main-stream work before the tail has run
            for t in (regression.values() if isinstance(regression, dict) else
                      regression if isinstance(regression, (list, tuple)) else [regression]):
                t.record_stream(self._side)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            if defer:   # loss-sum exchange + finalize: nothing on the main stream waits for them
                self._finish_loss(stats_now)   # this step's ring vector: not reused before the tail is waited for
            self.detect_tail(regression)
            self._side_done = torch.cuda.Event()
            self._side_done.record(self._side)
        return self.loss_out, self.det_count

    def capture(self, gt_boxes, gt_classes, gt_counts, logits, regression, fused=True, warmup=2, with_flip=False):
        """Record one whole `step()` on these (static) buffers into a CUDA graph and return it;
        `graph.replay()` then re-runs the step with one launch -- for small workloads (Shapes, VOC, small
        batches) the ~10 kernel launches of a step cost more than the kernels.  Refill the same input
        tensors in place between replays; outputs land in the usual members (loss_out, dlogits, det_*).
        Multi-rank: works with the peer-memory exchange (its epochs live in device memory and advance inside the
        kernel); an NCCL collective is not captured here."""
        if self._multi_rank():
            self._allreduce('counts', torch.zeros_like(self.stats))   # builds the peer exchange (collective set-up)
            if self._peer is None or self.exchange == 'nccl':
                raise RuntimeError('DenseHeadPath.capture with several ranks needs the peer-memory exchange '
                                   "(exchange='peer'); the NCCL fallback is not captured")
        self.join()
        self._static_stats = True   # one fixed statistics vector, zeroed by an op inside the graph
        hook, self.stage_hook = self.stage_hook, None
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):     # warm-up outside the capture: allocates dlogits / dreg
                for _ in range(max(1, warmup)):
                    self.step(gt_boxes, gt_classes, gt_counts, logits, regression, fused=fused, with_flip=with_flip)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.step(gt_boxes, gt_classes, gt_counts, logits, regression, fused=fused, with_flip=with_flip)
        finally:
            self.stage_hook = hook
        return graph

    def join(self):
        """Make the current stream wait for an overlapped detection tail."""
        if self._side_done is not None:
            torch.cuda.current_stream().wait_event(self._side_done)
            self._side_done = None

    def check_status(self, ignore=0):
        """Synchronising: raises what the reference's runtime asserts would have raised."""
        self.join()
        word = int(self.status.item())
        self.status.zero_()
        raise_for_status(word, ignore)
