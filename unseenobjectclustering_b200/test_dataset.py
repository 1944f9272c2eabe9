"""Host-side mirror of the reference's inference driver lib/fcn/test_dataset.py (same function
names, argument meaning and return types/devices), with everything kept on the GPU until the
API boundary.  clustering_features -> one batched uoc_meanshift_cluster call; the two-stage
plumbing (filter_labels_depth, crop_rois, match_label_crop) runs in csrc/refine.cu (a few launches
per frame for all objects, one host round trip for the number of boxes).  There is no CPU path: the
tensors that carry the pixels (depth, rgb, crop masks) must live on the GPU, label maps may come from
the CPU (the reference returns out_label on the CPU) and are uploaded.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import mean_shift as _ms

_refine_ws = {}


def _refine_workspace(dev, N, K):
    lib = _lib.load()
    nbytes = int(lib.uoc_refine_workspace_bytes(int(N), int(K)))
    sid = torch.cuda.current_stream(dev).cuda_stream
    key = (dev.index, sid)
    ws = _refine_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        _refine_ws[key] = ws
    return ws

CROP_SIZE = 224           # cfg.TRAIN.SYN_CROP_SIZE, lib/fcn/config.py:129
PADDING_PERCENTAGE = 0.25  # lib/fcn/test_dataset.py:66
METRIC = 'cosine'         # cfg.TRAIN.EMBEDDING_METRIC as read by clustering_features (lib/fcn/test_dataset.py:45);
                          # every shipped yml says cosine
_LIVE_CFG = [None]        # a reference cfg object read at call time instead (set by shim.install)


def _metric(metric=None):
    if metric is not None:
        return metric
    if _LIVE_CFG[0] is not None:
        return str(_LIVE_CFG[0].TRAIN.EMBEDDING_METRIC)
    return METRIC



def clustering_features(features, num_seeds=100, first_indices=None, flags=0, metric=None):
    """lib/fcn/test_dataset.py:44-59.  Returns (out_label float32 CPU [N,H,W], list of N int64 CPU
    [num_seeds] tensors).  One np.random.randint(0, n) is consumed per item, in order, exactly like
    the reference (lib/utils/mean_shift.py:155).  metric None -> the module-level METRIC."""
    N, _, H, W = features.shape
    labels_f32 = torch.empty((N, H, W), dtype=torch.float32, device=features.device) if features.is_cuda else None
    labels, selected = clustering_features_device(features, num_seeds, first_indices, flags, metric, labels_f32_out=labels_f32)
    out_label = labels_f32.cpu()                    # the reference's out_label: float32 on the CPU (test_dataset.py:48,57)
    sel = selected.cpu()
    _lib.raise_on_device_error(features.device)      # the host has just synchronised: surface kernel time-outs here
    return out_label, [sel[j] for j in range(N)]


def clustering_features_device(features, num_seeds=100, first_indices=None, flags=0, metric=None, labels_f32_out=None):
    """Same, but results stay on the device: (int32 [N, H*W], int64 [N, num_seeds])."""
    return _ms.cluster_fields(features, num_seeds=num_seeds, kappa=20.0, max_iters=10, first_indices=first_indices,
                              flags=flags, metric=_metric(metric), labels_f32_out=labels_f32_out)


# --------------------------------------------------------------------------------------------
# depth filter (lib/fcn/test_dataset.py:183-198)
# --------------------------------------------------------------------------------------------
def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.UocError("%s must be a CUDA tensor: there is no CPU path in this package" % what)


def _filter_labels_depth_device(labels, depth, threshold):
    """labels int [N,H,W] (device), depth [N,3,H,W].  Zero ids whose valid-depth fraction < threshold."""
    _require_cuda(depth, "depth")
    N = labels.shape[0]
    lib = _lib.load()
    dev = depth.device
    H, W = labels.shape[1], labels.shape[2]
    lab = labels.to(device=dev, dtype=torch.int32).contiguous()
    dep = depth.to(dtype=torch.float32).contiguous()
    out = torch.empty_like(lab)
    with torch.cuda.device(dev):
        ws = _refine_workspace(dev, N, 1)
        zptr = ctypes.c_void_p(dep.data_ptr() + 2 * H * W * 4)
        _lib.check(lib.uoc_filter_labels_depth(_lib.ptr(lab), zptr, 3 * H * W, N, H * W, float(threshold), _lib.ptr(out),
                                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "uoc_filter_labels_depth")
    return out


def filter_labels_depth(labels, depth, threshold):
    """Reference signature: labels float32 [N,H,W] (CPU or device), depth [N,3,H,W]; returns a new
    tensor on labels' device."""
    out = _filter_labels_depth_device(labels, depth, threshold)
    return out.to(labels.dtype).to(labels.device)


# --------------------------------------------------------------------------------------------
# ROI crops (lib/fcn/test_dataset.py:62-112, lib/utils/mask.py:180-195)
# --------------------------------------------------------------------------------------------
def crop_rois(rgb, initial_masks, depth, crop_size=CROP_SIZE):
    """lib/fcn/test_dataset.py:62-112.  Only batch item 0 is cropped, like the reference.
    Returns (rgb_crops [K,3,S,S], mask_crops [K,S,S], rois [K,4] float, depth_crops | None) on
    rgb's device."""
    _require_cuda(rgb, "rgb")
    dev = rgb.device
    lib = _lib.load()
    H, W = int(initial_masks.shape[1]), int(initial_masks.shape[2])
    lab0 = initial_masks[0].to(device=dev, dtype=torch.int32).contiguous()
    rgb0 = rgb[0].to(torch.float32).contiguous()
    dep0 = depth[0].to(device=dev, dtype=torch.float32).contiguous() if depth is not None else None
    with torch.cuda.device(dev):
        ws = _refine_workspace(dev, 1, 1)
        count_ids = torch.empty((257,), dtype=torch.int32, device=dev)
        rois_all = torch.empty((256, 4), dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr(dev)
        _lib.check(lib.uoc_crop_boxes(_lib.ptr(lab0), H, W, float(PADDING_PERCENTAGE), _lib.ptr(count_ids),
                                      _lib.ptr(rois_all), _lib.ptr(ws), ws.numel(), sp), "uoc_crop_boxes")
        K = int(count_ids[0].item())                    # the one host round trip: the output shapes depend on it
        _lib.raise_on_device_error(dev)
        rgb_crops = torch.empty((K, 3, crop_size, crop_size), dtype=torch.float32, device=dev)
        mask_crops = torch.empty((K, crop_size, crop_size), dtype=torch.float32, device=dev)
        depth_crops = torch.empty((K, 3, crop_size, crop_size), dtype=torch.float32, device=dev) if dep0 is not None else None
        if K > 0:
            ids_ptr = ctypes.c_void_p(count_ids.data_ptr() + 4)
            _lib.check(lib.uoc_crop_resize(_lib.ptr(rgb0), _lib.ptr(dep0), _lib.ptr(lab0), H, W, ids_ptr, _lib.ptr(rois_all), K,
                                           int(crop_size), _lib.ptr(rgb_crops), _lib.ptr(mask_crops), _lib.ptr(depth_crops), sp),
                       "uoc_crop_resize")
    return rgb_crops, mask_crops, rois_all[:K].clone(), depth_crops


# --------------------------------------------------------------------------------------------
# merge crop labels back (lib/fcn/test_dataset.py:116-179)
# --------------------------------------------------------------------------------------------
def match_label_crop(initial_masks, labels_crop, out_label_crop, rois, depth_crop):
    """Returns (refined_masks float32 [N,H,W] on the CPU like the reference's `refined_masks`,
    labels_crop with dropped clusters set to -1)."""
    _require_cuda(out_label_crop, "out_label_crop")
    dev = out_label_crop.device
    lib = _lib.load()
    K, S = int(labels_crop.shape[0]), int(labels_crop.shape[1])
    H, W = int(initial_masks.shape[-2]), int(initial_masks.shape[-1])
    lc32 = labels_crop.to(device=dev, dtype=torch.int32).contiguous()
    mc = out_label_crop.to(torch.float32).contiguous()
    ro = rois.to(device=dev, dtype=torch.float32).contiguous()
    dc = depth_crop.to(torch.float32).contiguous() if depth_crop is not None else None
    refined = torch.zeros(tuple(initial_masks.shape), dtype=torch.float32, device=dev)
    lc_out = torch.empty_like(lc32)
    with torch.cuda.device(dev):
        ws = _refine_workspace(dev, 1, max(K, 1))
        # only batch item 0 is refined, like the reference (test_dataset.py:177)
        _lib.check(lib.uoc_match_label_crop(_lib.ptr(lc32), _lib.ptr(mc), _lib.ptr(ro), _lib.ptr(dc), K, S, H, W,
                                            _lib.ptr(refined), _lib.ptr(lc_out), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr(dev)), "uoc_match_label_crop")
    refined_cpu = refined.cpu()
    _lib.raise_on_device_error(dev)
    return refined_cpu, lc_out.to(labels_crop.dtype)


# --------------------------------------------------------------------------------------------
# single frame (lib/fcn/test_dataset.py:232-267)
# --------------------------------------------------------------------------------------------
def _uses_depth(network):
    """lib/fcn/test_dataset.py:236-239: depth is used iff cfg.INPUT is 'DEPTH' or 'RGBD' -- read from the network this
    package built (its input_type mirrors cfg.INPUT at construction), else from the live reference cfg, else RGBD."""
    net = getattr(network, "module", network)            # torch.nn.DataParallel wrapper
    it = getattr(net, "input_type", None)
    if it is None and _LIVE_CFG[0] is not None:
        it = str(_LIVE_CFG[0].INPUT)
    return it is None or it in ("DEPTH", "RGBD")


def test_sample(sample, network, network_crop, first_indices=None, first_indices_crop=None, flags=0):
    """Same contract as the reference: sample['image_color'] / ['depth'] are [1,3,H,W] float32
    tensors; returns (out_label, out_label_refined | None) as float32 CPU [N,H,W]."""
    image = sample['image_color'].cuda()
    depth = None
    if _uses_depth(network) and sample.get('depth') is not None:
        depth = sample['depth'].cuda()
    label = sample['label'].cuda() if 'label' in sample else None

    features = network(image, label, depth).detach()
    N, _, H, W = features.shape
    labels_f32 = torch.empty((N, H, W), dtype=torch.float32, device=features.device) if depth is None else None
    labels, _ = clustering_features_device(features, 100, first_indices, flags, labels_f32_out=labels_f32)
    labels = labels.view(N, H, W)
    if depth is not None:
        labels = _filter_labels_depth_device(labels, depth, 0.8)
        labels_f32 = labels.to(torch.float32)
    out_label = labels_f32.cpu()
    _lib.raise_on_device_error(image.device)

    out_label_refined = None
    if network_crop is not None:
        rgb_crop, out_label_crop, rois, depth_crop = crop_rois(image, labels, depth)
        if rgb_crop.shape[0] > 0:
            features_crop = network_crop(rgb_crop, out_label_crop, depth_crop).detach()
            labels_crop, _ = clustering_features_device(features_crop, 100, first_indices_crop, flags)
            K = rgb_crop.shape[0]
            labels_crop = labels_crop.view(K, CROP_SIZE, CROP_SIZE)
            out_label_refined, _ = match_label_crop(out_label, labels_crop, out_label_crop, rois, depth_crop)
    return out_label, out_label_refined


class SegmentationClustering(object):
    """Convenience wrapper named in BASELINE.json's north_star: network(s) + two-stage clustering."""

    def __init__(self, network, network_crop=None, flags=0):
        self.network = network
        self.network_crop = network_crop
        self.flags = flags

    def __call__(self, image, depth, first_indices=None, first_indices_crop=None):
        return test_sample({'image_color': image, 'depth': depth}, self.network, self.network_crop, first_indices,
                           first_indices_crop, self.flags)
