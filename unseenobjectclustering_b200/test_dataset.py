"""Host-side mirror of the reference's inference driver lib/fcn/test_dataset.py (same function
names, argument meaning and return types/devices), with everything kept on the GPU until the
API boundary.  clustering_features -> one batched uoc_meanshift_cluster call; the two-stage
plumbing (filter_labels_depth, crop_rois, match_label_crop) runs in csrc/refine.cu for CUDA tensors
(a few launches per frame for all objects, one host round trip for the number of boxes); for CPU
tensors the same functions are plain torch tensor plumbing (host logic tests).
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

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
    labels, selected = clustering_features_device(features, num_seeds, first_indices, flags, metric)
    N, _, H, W = features.shape
    out_label = labels.view(N, H, W).to(torch.float32).cpu()
    sel = selected.cpu()
    return out_label, [sel[j] for j in range(N)]


def clustering_features_device(features, num_seeds=100, first_indices=None, flags=0, metric=None):
    """Same, but results stay on the device: (int32 [N, H*W], int64 [N, num_seeds])."""
    return _ms.cluster_fields(features, num_seeds=num_seeds, kappa=20.0, max_iters=10, first_indices=first_indices,
                              flags=flags, metric=_metric(metric))


# --------------------------------------------------------------------------------------------
# depth filter (lib/fcn/test_dataset.py:183-198)
# --------------------------------------------------------------------------------------------
def _filter_labels_depth_device(labels, depth, threshold, max_label=256):
    """labels int [N,H,W] (device), depth [N,3,H,W].  Zero ids whose valid-depth fraction < threshold."""
    N = labels.shape[0]
    if labels.is_cuda:
        lib = _lib.load()
        dev = labels.device
        H, W = labels.shape[1], labels.shape[2]
        lab = labels.to(torch.int32).contiguous()
        dep = depth.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty_like(lab)
        with torch.cuda.device(dev):
            ws = _refine_workspace(dev, N, 1)
            zptr = ctypes.c_void_p(dep.data_ptr() + 2 * H * W * 4)
            _lib.check(lib.uoc_filter_labels_depth(_lib.ptr(lab), zptr, 3 * H * W, N, H * W, float(threshold), _lib.ptr(out),
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "uoc_filter_labels_depth")
        return out.to(labels.dtype)
    lab = labels.reshape(N, -1).to(torch.int64)
    valid = (depth[:, 2].reshape(N, -1) > 0).to(torch.float32)
    tot = torch.zeros((N, max_label), dtype=torch.float32, device=lab.device).scatter_add_(1, lab, torch.ones_like(valid))
    good = torch.zeros((N, max_label), dtype=torch.float32, device=lab.device).scatter_add_(1, lab, valid)
    frac = good / tot.clamp(min=1.0)
    drop = (frac < threshold) & (tot > 0)
    drop[:, 0] = False
    out = torch.where(torch.gather(drop, 1, lab), torch.zeros_like(lab), lab)
    return out.view_as(labels).to(labels.dtype)


def filter_labels_depth(labels, depth, threshold):
    """Reference signature: labels float32 [N,H,W] (CPU or device), depth [N,3,H,W]; returns a new
    tensor on labels' device."""
    dev = depth.device
    out = _filter_labels_depth_device(labels.to(dev).to(torch.int64), depth, threshold)
    return out.to(labels.dtype).to(labels.device)


# --------------------------------------------------------------------------------------------
# ROI crops (lib/fcn/test_dataset.py:62-112, lib/utils/mask.py:180-195)
# --------------------------------------------------------------------------------------------
def _rois_from_labels(label0, max_label=256):
    """label0: int64 [H,W] on device.  Returns (ids int64 [K], rois int64 [K,4] x0,y0,x1,y1 padded
    and clamped) with ONE host sync."""
    H, W = label0.shape
    flat = label0.reshape(-1)
    ys = torch.arange(H, device=flat.device).view(H, 1).expand(H, W).reshape(-1)
    xs = torch.arange(W, device=flat.device).view(1, W).expand(H, W).reshape(-1)
    big = torch.full((max_label,), 1 << 30, dtype=torch.int64, device=flat.device)
    small = torch.full((max_label,), -1, dtype=torch.int64, device=flat.device)
    x0 = big.clone().scatter_reduce_(0, flat, xs, reduce="amin")
    y0 = big.clone().scatter_reduce_(0, flat, ys, reduce="amin")
    x1 = small.clone().scatter_reduce_(0, flat, xs, reduce="amax")
    y1 = small.clone().scatter_reduce_(0, flat, ys, reduce="amax")
    box = torch.stack([x0, y0, x1, y1], 1).cpu().numpy()
    ids, rois = [], []
    for lid in range(1, max_label):                      # ascending ids == torch.unique order, 0 skipped
        if box[lid, 2] < 0:
            continue
        bx0, by0, bx1, by1 = (int(v) for v in box[lid])
        xp = int(np.round(np.float32(bx1 - bx0) * np.float32(PADDING_PERCENTAGE)))   # half-to-even like torch.round
        yp = int(np.round(np.float32(by1 - by0) * np.float32(PADDING_PERCENTAGE)))
        rois.append([max(bx0 - xp, 0), max(by0 - yp, 0), min(bx1 + xp, W - 1), min(by1 + yp, H - 1)])
        ids.append(lid)
    return ids, rois


def crop_rois(rgb, initial_masks, depth, crop_size=CROP_SIZE):
    """lib/fcn/test_dataset.py:62-112.  Only batch item 0 is cropped, like the reference.
    Returns (rgb_crops [K,3,S,S], mask_crops [K,S,S], rois [K,4] float, depth_crops | None) on
    rgb's device."""
    dev = rgb.device
    if rgb.is_cuda:
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
            rgb_crops = torch.empty((K, 3, crop_size, crop_size), dtype=torch.float32, device=dev)
            mask_crops = torch.empty((K, crop_size, crop_size), dtype=torch.float32, device=dev)
            depth_crops = torch.empty((K, 3, crop_size, crop_size), dtype=torch.float32, device=dev) if dep0 is not None else None
            if K > 0:
                ids_ptr = ctypes.c_void_p(count_ids.data_ptr() + 4)
                _lib.check(lib.uoc_crop_resize(_lib.ptr(rgb0), _lib.ptr(dep0), _lib.ptr(lab0), H, W, ids_ptr, _lib.ptr(rois_all), K,
                                               int(crop_size), _lib.ptr(rgb_crops), _lib.ptr(mask_crops), _lib.ptr(depth_crops), sp),
                           "uoc_crop_resize")
        return rgb_crops, mask_crops, rois_all[:K].clone(), depth_crops
    masks0 = initial_masks[0].to(dev).to(torch.int64)
    ids, rois_l = _rois_from_labels(masks0)
    K = len(ids)
    size = (crop_size, crop_size)
    rgb_crops = torch.zeros((K, 3, crop_size, crop_size), device=dev)
    mask_crops = torch.zeros((K, crop_size, crop_size), device=dev)
    depth_crops = torch.zeros((K, 3, crop_size, crop_size), device=dev) if depth is not None else None
    rois = torch.tensor(rois_l, dtype=torch.float32, device=dev).view(K, 4)
    for k, (lid, (x0, y0, x1, y1)) in enumerate(zip(ids, rois_l)):
        rgb_crops[k] = F.interpolate(rgb[0:1, :, y0:y1 + 1, x0:x1 + 1], size=size, mode="bilinear", align_corners=True)[0]
        m = (masks0[y0:y1 + 1, x0:x1 + 1] == lid).to(torch.float32)
        mask_crops[k] = F.interpolate(m[None, None], size=size, mode="nearest")[0, 0]
        if depth is not None:
            depth_crops[k] = F.interpolate(depth[0:1, :, y0:y1 + 1, x0:x1 + 1], size=size, mode="bilinear",
                                           align_corners=True)[0]
    return rgb_crops, mask_crops, rois, depth_crops


# --------------------------------------------------------------------------------------------
# merge crop labels back (lib/fcn/test_dataset.py:116-179)
# --------------------------------------------------------------------------------------------
def match_label_crop(initial_masks, labels_crop, out_label_crop, rois, depth_crop, max_label=256):
    """Returns (refined_masks float32 [N,H,W] on the CPU like the reference's `refined_masks`,
    labels_crop with dropped clusters set to -1)."""
    dev = out_label_crop.device
    if out_label_crop.is_cuda:
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
        return refined.cpu(), lc_out.to(labels_crop.dtype)
    lc = labels_crop.to(dev).to(torch.int64)
    K = lc.shape[0]
    flat = lc.view(K, -1)
    # (i) clusters overlapping the stage-1 mask by < 50 % of their own area -> -1
    area = torch.zeros((K, max_label), device=dev).scatter_add_(1, flat, torch.ones(flat.shape, device=dev))
    inter = torch.zeros((K, max_label), device=dev).scatter_add_(1, flat, out_label_crop.view(K, -1).to(torch.float32))
    drop = (inter / area.clamp(min=1.0) < 0.5) & (area > 0)
    lc = torch.where(torch.gather(drop, 1, flat), torch.full_like(flat, -1), flat).view_as(lc)
    # (ii) far -> near ordering
    if depth_crop is not None:
        z = depth_crop[:, 2]
        kept = lc > -1
        has_kept = kept.view(K, -1).any(1)
        use = torch.where(has_kept.view(K, 1, 1), kept, torch.ones_like(kept)) & (z > 0)
        keys = (z * use).view(K, -1).sum(1) / use.view(K, -1).sum(1)     # NaN when empty, like torch.mean([])
    else:
        keys = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 2] - rois[:, 0] + 1)
    order = torch.argsort(keys, descending=True, stable=True).cpu().tolist()
    rois_h = rois.cpu().numpy().astype(np.int64)
    present = (torch.zeros((K, max_label + 1), device=dev)
               .scatter_add_(1, (lc.view(K, -1) + 1), torch.ones((K, lc[0].numel()), device=dev)) > 0).cpu().numpy()
    # (iii) renumber 1,2,3.. in that order, paste nearest-resized non-zeros
    refined = torch.zeros(tuple(initial_masks.shape), dtype=torch.float32, device=dev)
    count = 0
    for i in order:
        lut = torch.zeros((max_label + 1,), dtype=torch.float32, device=dev)
        for lid in range(max_label):
            if present[i, lid + 1]:
                count += 1
                lut[lid + 1] = count
        renum = lut[lc[i] + 1]
        x0, y0, x1, y1 = (int(v) for v in rois_h[i])
        back = F.interpolate(renum[None, None], size=(y1 - y0 + 1, x1 - x0 + 1), mode="nearest")[0, 0]
        region = refined[0, y0:y1 + 1, x0:x1 + 1]
        refined[0, y0:y1 + 1, x0:x1 + 1] = torch.where(back != 0, back, region)
    return refined.cpu(), lc.to(labels_crop.dtype)


# --------------------------------------------------------------------------------------------
# single frame (lib/fcn/test_dataset.py:232-267)
# --------------------------------------------------------------------------------------------
def test_sample(sample, network, network_crop, first_indices=None, first_indices_crop=None, flags=0):
    """Same contract as the reference: sample['image_color'] / ['depth'] are [1,3,H,W] float32
    tensors; returns (out_label, out_label_refined | None) as float32 CPU [N,H,W]."""
    image = sample['image_color'].cuda()
    depth = sample['depth'].cuda() if 'depth' in sample and sample['depth'] is not None else None
    label = sample['label'].cuda() if 'label' in sample else None

    features = network(image, label, depth).detach()
    labels, _ = clustering_features_device(features, 100, first_indices, flags)
    N, _, H, W = features.shape
    labels = labels.view(N, H, W)
    if depth is not None:
        labels = _filter_labels_depth_device(labels, depth, 0.8)
    out_label = labels.to(torch.float32).cpu()

    out_label_refined = None
    if network_crop is not None:
        rgb_crop, out_label_crop, rois, depth_crop = crop_rois(image, labels, depth)
        if rgb_crop.shape[0] > 0:
            features_crop = network_crop(rgb_crop, out_label_crop, depth_crop).detach()
            labels_crop, _ = clustering_features_device(features_crop, 100, first_indices_crop, flags)
            K = rgb_crop.shape[0]
            labels_crop = labels_crop.view(K, CROP_SIZE, CROP_SIZE).to(torch.float32)
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
