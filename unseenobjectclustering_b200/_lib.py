"""ctypes binding of libuoc_b200.so (the C ABI declared in include/uoc.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing or a call fails, the
caller gets an exception carrying uoc_last_error().
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libuoc_b200.so")

FLAG_LOOP_SIMT = 1
FLAG_CONV_SIMT = 2
FLAG_SYNC_CHECK = 4
FLAG_FPS_FP32 = 8
FLAG_EUCLIDEAN = 16
MAX_SEEDS = 128


class UocError(RuntimeError):
    pass


class WeightDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data_host", ctypes.c_void_p), ("numel", ctypes.c_int64)]


_c = ctypes
_vp, _i, _i64, _f, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); every symbol include/uoc.h declares
SIGNATURES = {
    "uoc_last_error": (_c.c_char_p, []),
    "uoc_version": (_i, []),
    "uoc_launch_count": (_c.c_uint64, []),
    "uoc_device_info": (_i, [_c.POINTER(_i), _c.POINTER(_i), _c.POINTER(_i)]),
    "uoc_set_knob": (_i, [_c.c_char_p, _i]),
    "uoc_check_device_error": (_i, [_vp]),
    "uoc_peek_device_error_async": (_i, [_vp, _vp]),
    "uoc_meanshift_workspace_bytes": (_sz, [_i, _i64, _i, _i]),
    "uoc_meanshift_cluster": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                   _i, _vp]),
    "uoc_meanshift_cluster_ex": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _sz, _i, _vp]),
    "uoc_assign_labels_typed": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "uoc_select_seeds": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "uoc_select_seeds_init": (_i, [_vp, _i64, _i64, _i, _i64, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "uoc_hill_climb": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _f, _i, _vp, _vp, _sz, _i, _vp]),
    "uoc_label_seeds": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    "uoc_assign_labels": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "uoc_label_seeds_ex": (_i, [_vp, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    "uoc_assign_labels_ex": (_i, [_vp, _i64, _i64, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "uoc_pack_bf16": (_i, [_vp, _i64, _i64, _i, _i64, _i, _vp, _vp]),
    "uoc_backbone_create": (_i, [_c.POINTER(_vp), _c.POINTER(WeightDesc), _i, _i]),
    "uoc_backbone_create_ex": (_i, [_c.POINTER(_vp), _c.POINTER(WeightDesc), _i, _i, _i, _i, _i]),
    "uoc_backbone_feature_dim": (_i, [_vp]),
    "uoc_backbone_destroy": (None, [_vp]),
    "uoc_backbone_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "uoc_backbone_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "uoc_backbone_read_trunk": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "uoc_conv2d_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "uoc_refine_workspace_bytes": (_sz, [_i, _i]),
    "uoc_filter_labels_depth": (_i, [_vp, _vp, _i64, _i, _i64, _f, _vp, _vp, _sz, _vp]),
    "uoc_crop_boxes": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "uoc_crop_resize": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "uoc_match_label_crop": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "uoc_metrics_workspace_bytes": (_sz, [_i, _i]),
    "uoc_multilabel_counts": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "uoc_prepare_inputs": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _f, _vp, _vp, _vp]),
    "uoc_compute_xyz": (_i, [_vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp]),
}

_lib = None


def load():
    """Load the library (once).  Raises UocError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UocError("%s not found: build it with `python -m unseenobjectclustering_b200.build` "
                       "(or __graft_entry__.build()); there is no fallback path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    msg = load().uoc_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(status, what):
    if status != 0:
        raise UocError("%s failed (status %d): %s" % (what, status, last_error()))


def device_info():
    lib = load()
    a, b, c = _i(0), _i(0), _i(0)
    check(lib.uoc_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "uoc_device_info")
    return a.value, b.value, c.value


def raise_on_device_error(device=None):
    """Read (and clear) the device error word of `device` on its current stream; raises UocError if a kernel reported a
    pipeline time-out or a bad configuration.  Synchronises the stream: call it where the host has just synchronised
    anyway (a .cpu() / .item()), never on the hot path."""
    import torch
    with torch.cuda.device(device):
        check(load().uoc_check_device_error(stream_ptr(device)), "device error check")


KNOB_DEFAULTS = {"conv_pair": -1, "conv_wres": -1, "conv_debug": 0, "conv_trace": 0, "fps_tc": 1, "fps_stream": 0, "fps_tmem_tiles": -1,
                 "fps_batch_stream": 0, "fps_rn_margin": 0, "fps_stats": 0, "loop_trace": 0, "assign_simt": 0}


_knob_epoch = 0


def set_knob(name, value):
    """Parity-test / measurement switch of the library (include/uoc.h uoc_set_knob); not part of the product contract."""
    global _knob_epoch
    check(load().uoc_set_knob(name.encode(), int(value)), "uoc_set_knob(%s)" % name)
    _knob_epoch += 1


def knob_epoch():
    """Changes whenever a knob is set: captured launch sequences (networks._GraphedForward) are keyed on it."""
    return _knob_epoch


def ptr(t):
    """Raw device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
