"""ctypes binding of librndense.so (include/rndense.h).

There is no CPU fallback: importing this module without the built library raises, and any
compute call without a CUDA device raises.  Build with `python -c "import __graft_entry__ as g;
g.build()"` or `make -C retinanet_tensorflow_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import threading
import os

RN_MAX_LEVELS = 8
RN_MAX_ANCHORS = 16
RN_STATS_HEADER = 8

RN_SIZE_TF1_INT_TRUNC = 0
RN_SIZE_FLOAT = 1
RN_LABEL_BACKGROUND = -1
RN_LABEL_IGNORE = -2

RN_CLS_NONE, RN_CLS_FOCAL, RN_CLS_BCE, RN_CLS_BALANCED_BCE = 0, 1, 2, 3
RN_CLS_DICE_FLAG = 16
RN_CLS_GENERIC_FLAG = 32
RN_DTYPE_F32, RN_DTYPE_BF16 = 0, 1
RN_SCORES_LOGITS, RN_SCORES_BF16 = 1, 2

RN_STAT_TRAINABLE, RN_STAT_NUM_FG, RN_STAT_CLS_SUM, RN_STAT_HUBER_SUM = 0, 1, 2, 3

RN_BOXOP_DECODE, RN_BOXOP_ADD_CENTRE, RN_BOXOP_SCALE = 0, 1, 2
RN_BOXOP_CENTRE_TO_CORNER, RN_BOXOP_CORNER_TO_CENTRE = 3, 4

RN_FLAG_BOX_ORDER, RN_FLAG_IOU_RANGE, RN_FLAG_CAND_OVERFLOW, RN_FLAG_DET_OVERFLOW, RN_FLAG_EMPTY_GT = 1, 2, 4, 8, 16
RN_FLAG_EXCHANGE_TIMEOUT = 32


class Level(C.Structure):
    _fields_ = [
        ('grid_h', C.c_int32), ('grid_w', C.c_int32), ('num_anchors', C.c_int32),
        ('tiles_x', C.c_int32), ('tiles_y', C.c_int32), ('tile_start', C.c_int32),
        ('anchor_offset', C.c_int64),
        ('y0', C.c_float), ('y_step', C.c_float), ('x0', C.c_float), ('x_step', C.c_float),
        ('max_half_h', C.c_float), ('max_half_w', C.c_float),
        ('anchor_h', C.c_float * RN_MAX_ANCHORS), ('anchor_w', C.c_float * RN_MAX_ANCHORS),
    ]


class Pyramid(C.Structure):
    _fields_ = [
        ('num_levels', C.c_int32), ('image_h', C.c_int32), ('image_w', C.c_int32),
        ('tiles_per_image', C.c_int32), ('anchors_per_image', C.c_int64),
        ('level', Level * RN_MAX_LEVELS),
    ]


class Rows(C.Structure):
    _fields_ = [('num_segments', C.c_int32), ('images', C.c_int32), ('rows', C.c_int64 * RN_MAX_LEVELS)]


class Maps(C.Structure):
    _fields_ = [('p', C.c_void_p * RN_MAX_LEVELS)]


_HERE = os.path.dirname(os.path.abspath(__file__))
# RNDENSE_LIB points at another build of the same library (tuning variants); default: the in-tree build
LIB_PATH = os.environ.get('RNDENSE_LIB') or os.path.join(_HERE, 'lib', 'librndense.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'librndense.so is not built (%s). This package has no CPU or PyTorch fallback: run '
        '`python -c "import __graft_entry__ as g; g.build()"` from the repo root.' % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_PP, _RP, _MP = C.POINTER(Pyramid), C.POINTER(Rows), C.POINTER(Maps)

# name -> (restype, argtypes); must list every RN_API symbol of include/rndense.h
PROTOTYPES = {
    'rn_version': (_i, []),
    'rn_last_error': (C.c_char_p, []),
    'rn_status_string': (C.c_char_p, [_i]),
    'rn_launch_count': (_i64, []),
    'rn_set_device': (_i, [_i]),
    'rn_set_option': (_i, [C.c_char_p, _i]),
    'rn_pyramid_build': (_i, [_i, _i, _i, _vp, _vp, _i, _i, _PP]),
    'rn_pyramid_from_grids': (_i, [_i, _vp, _vp, _i, _PP]),
    'rn_rows_from_pyramid': (_i, [_PP, _i, _RP]),
    'rn_maps_from_merged': (_i, [_PP, _i, _vp, _sz, _MP]),
    'rn_anchor_generate': (_i, [_PP, _vp, _vp, _vp]),
    'rn_assign_encode': (_i, [_PP, _i, _vp, _vp, _vp, _i, _i, _f, _f, _MP, _MP, _MP, _MP, _vp, _vp, _vp]),
    'rn_assign_encode_patch': (_i, [_PP, _i, _vp, _vp, _vp, _i, _i, _f, _f, _MP, _MP, _MP, _MP, _vp, _vp, _i, _i, _vp]),
    'rn_labels_dense': (_i, [_RP, _i, _MP, _MP, _MP, _vp]),
    'rn_flip_labels': (_i, [_PP, _i, _MP, _MP, _MP, _MP, _vp]),
    'rn_iou_pairwise': (_i, [_i64, _vp, _vp, _vp, _vp, _vp]),
    'rn_label_stats': (_i, [_RP, _i, _MP, _vp, _vp]),
    'rn_label_stats_dense': (_i, [_i64, _i, _vp, _vp, _vp]),
    'rn_loss_workspace_bytes': (_sz, [_i]),
    'rn_cls_sums': (_i, [_RP, _i, _MP, _MP, _i, _i, _vp, _vp, _sz, _vp]),
    'rn_cls_loss_fwd_bwd': (_i, [_RP, _i, _MP, _MP, _i, _i, _f, _f, _f, _MP, _i, _vp, _vp, _sz, _vp]),
    'rn_huber_fwd_bwd': (_i, [_RP, _MP, _MP, _MP, _MP, _MP, _vp, _vp, _sz, _vp]),
    'rn_head_loss': (_i, [_RP, _i, _MP, _MP, _i, _f, _f, _f, _MP, _i, _MP, _MP, _MP, _vp, _vp, _vp, _sz, _vp, _f, _i, _vp, _vp,
                     _vp, _sz, _vp]),
    'rn_huber_rows': (_i, [_i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'rn_loss_finalize': (_i, [_i, _i, _vp, _vp, _vp]),
    'rn_cls_loss_elementwise': (_i, [_i64, _i, _vp, _vp, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    'rn_focal_softmax': (_i, [_i64, _i, _vp, _vp, _f, _f, _f, _vp, _vp, _vp]),
    'rn_colsum_backward': (_i, [_i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'rn_sigmoid': (_i, [_i64, _vp, _vp, _vp]),
    'rn_mask_workspace_bytes': (_sz, [_RP]),
    'rn_mask_scan': (_i, [_RP, _MP, _vp, _sz, _vp, _vp]),
    'rn_mask_gather': (_i, [_RP, _i, _MP, _MP, _vp, _vp, _vp]),
    'rn_shapes_render': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'rn_draw_boxes': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'rn_stats_mailbox_bytes': (_sz, [_i, _i, _i]),
    'rn_stats_exchange': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, C.c_uint64, _vp, _vp]),
    'rn_stats_exchange_auto': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'rn_decode_boxes': (_i, [_PP, _i, _MP, _MP, _vp]),
    'rn_boxmap_op': (_i, [_PP, _i, _i, _MP, _MP, _vp]),
    'rn_classmap_fg': (_i, [_i64, _i, _vp, _f, _vp, _vp]),
    'rn_select_workspace_bytes': (_sz, [_i, _i]),
    'rn_decode_select': (_i, [_PP, _i, _i, _MP, _i, _MP, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'rn_select_finalize': (_i, [_PP, _i, _MP, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'rn_fused_select_supported': (_i, [_i, _i, _f, _RP]),
    'rn_cls_loss_select': (_i, [_RP, _i, _MP, _MP, _i, _f, _f, _f, _MP, _i, _vp, _vp, _sz, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    'rn_nms_workspace_bytes': (_sz, [_i, _i, _i]),
    'rn_nms_classwise': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'rn_detect_tail_supported': (_i, [_i, _i]),
    'rn_detect_tail': (_i, [_PP, _i, _i, _MP, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp,
                       _vp, _vp, _vp, _sz, _vp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the .so is stale: rebuild
    _fn.restype = _res
    _fn.argtypes = _args


class RnError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib.rn_last_error().decode('utf-8', 'replace')
        super().__init__('%s failed: %s (%s)' % (where, lib.rn_status_string(status).decode(), msg))


def check(status, where):
    if status != 0:
        if status in (-1, -2):
            raise ValueError(str(RnError(status, where)))
        raise RnError(status, where)


def call(name, *args):
    check(getattr(lib, name)(*args), name)


_device_set = {}
_tls = threading.local()


def require_cuda(device=None):
    """The product path runs on a B200 only.  Returns the torch device and makes it the calling thread's current
    CUDA device for the library (rn_set_device) -- once per thread and device, not per call."""
    import torch
    if device is None:
        cached = getattr(_tls, 'current', None)
        if cached is not None and cached[0] == torch.cuda.current_device():
            return cached[1]
    if not torch.cuda.is_available():
        raise RuntimeError('retinanet_tensorflow_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError('retinanet_tensorflow_b200 tensors must live on a CUDA device, got %s' % dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if not _device_set.get(idx):
        cap = torch.cuda.get_device_capability(idx)
        if cap[0] != 10:
            raise RuntimeError('librndense.so is compiled for sm_100a only; device %d is sm_%d%d' % (idx, *cap))
        _device_set[idx] = True
    call('rn_set_device', idx)
    dev = torch.device('cuda', idx)
    _tls.current = (idx, dev) if idx == torch.cuda.current_device() else None
    return dev


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return C.c_void_p(t.data_ptr())


def maps_of(tensors):
    """rn_maps_t from a list of per-level tensors (or None)."""
    m = Maps()
    if tensors is not None:
        for i, t in enumerate(tensors):
            assert t.is_cuda and t.is_contiguous(), 'expected contiguous CUDA tensors'
            m.p[i] = t.data_ptr()
    return m


def rows_of(row_counts, images=1):
    r = Rows()
    r.num_segments = len(row_counts)
    r.images = images
    for i, n in enumerate(row_counts):
        r.rows[i] = int(n)
    return r
