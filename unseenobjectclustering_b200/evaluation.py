"""Host-side mirror of the reference's lib/utils/evaluation.py:multilabel_metrics (the evaluation tail of
fcn.test_dataset.test_segnet, lib/fcn/test_dataset.py:307-330).

The pixel work -- true-positive counts of every (ground-truth, predicted) label pair, the one-pixel boundary maps of
every label (seg2bmap) and their disk-dilated overlaps for every pair (boundary_overlap) -- runs in csrc/metrics.cu in
three launches; the reference does it with two boolean masks, two boundary maps and two cv2.dilate calls PER PAIR on
the CPU.  The Hungarian matching of the (#gt x #pred) F-measure matrix and the final ratios stay on the host, in the
reference's float64 arithmetic.  Same signature, same dictionary.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

BACKGROUND_LABEL = 0            # lib/utils/evaluation.py:11
_KL = 256


def _assignment(cost):
    """Minimum-cost assignment of a rectangular matrix as a list of (row, col) pairs -- what
    utils.munkres.Munkres().compute returns (evaluation.py:221-223).  The optimum is the same; when several
    assignments are optimal the pairs may differ from the reference's implementation."""
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(cost)
    return list(zip(r.tolist(), c.tolist()))


def multilabel_counts(prediction, gt, device=None):
    """The device part: returns (tp [256,256] (gt label, predicted label), boundary precision true positives
    [256,256], boundary recall true positives [256,256], (boundary_prec_denom, boundary_rec_denom)) as numpy."""
    lib = _lib.load()
    if device is None:
        device = prediction.device if torch.is_tensor(prediction) and prediction.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)

    def as_labels(a):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        if t.dim() != 2:
            raise ValueError("prediction / gt must be [H, W] label maps")
        td = t.to(device=device)
        ti = td.to(torch.int32).contiguous()
        if not torch.equal(ti.to(torch.float64), td.to(torch.float64)):
            raise ValueError("label maps must hold integers")
        return ti

    p, g = as_labels(prediction), as_labels(gt)
    if p.shape != g.shape:
        raise ValueError("prediction and gt must have the same shape")
    if int(torch.minimum(p.min(), g.min())) < 0 or int(torch.maximum(p.max(), g.max())) > 254:
        raise _lib.UocError("label ids must be in [0, 254]")
    H, W = int(p.shape[0]), int(p.shape[1])
    # bound_pix = bound_th if bound_th >= 1 else ceil(bound_th * ||shape||)   (evaluation.py:85-86, bound_th = 0.003)
    bound_pix = int(np.ceil(0.003 * np.linalg.norm((H, W))))
    with torch.cuda.device(device):
        tp = torch.empty((_KL, _KL), dtype=torch.int32, device=device)
        bp = torch.empty((_KL, _KL), dtype=torch.int32, device=device)
        br = torch.empty((_KL, _KL), dtype=torch.int32, device=device)
        den = torch.empty((2,), dtype=torch.int64, device=device)
        ws = torch.empty(int(lib.uoc_metrics_workspace_bytes(H, W)) + 256, dtype=torch.uint8, device=device)
        _lib.check(lib.uoc_multilabel_counts(_lib.ptr(p), _lib.ptr(g), H, W, bound_pix, _lib.ptr(tp), _lib.ptr(bp), _lib.ptr(br),
                                             _lib.ptr(den), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device)),
                   "uoc_multilabel_counts")
        out = torch.stack([tp, bp, br]).cpu().numpy().astype(np.int64)
        d = den.cpu().numpy()
    return out[0], out[1], out[2], (int(d[0]), int(d[1]))


def multilabel_metrics(prediction, gt, obj_detect_threshold=0.75):
    """lib/utils/evaluation.py:109-257.  prediction, gt: [H, W] numpy arrays (or tensors) of label ids, 0 = background."""
    tp_all, bprec_all, brec_all, dens = multilabel_counts(prediction, gt)
    return metrics_from_counts(tp_all, bprec_all, brec_all, dens, obj_detect_threshold)


def metrics_from_counts(tp_all, bprec_all, brec_all, dens, obj_detect_threshold=0.75):
    """The host half of multilabel_metrics: from the [256,256] count tables (indexed [gt label][predicted label]) and the
    two boundary denominators to the reference's dictionary, in its float64 arithmetic (evaluation.py:139-257)."""
    bden_p, bden_g = dens
    area_gt = tp_all.sum(axis=1)                 # pixels per ground-truth label
    area_pred = tp_all.sum(axis=0)               # pixels per predicted label
    labels_gt = [l for l in range(1, _KL) if area_gt[l] > 0]
    labels_pred = [l for l in range(1, _KL) if area_pred[l] > 0]
    num_labels_gt, num_labels_pred = len(labels_gt), len(labels_pred)

    def edge(f, p, r, pct):
        return {'Objects F-measure': f, 'Objects Precision': p, 'Objects Recall': r,
                'Boundary F-measure': f, 'Boundary Precision': p, 'Boundary Recall': r,
                'obj_detected': num_labels_pred, 'obj_detected_075': 0., 'obj_gt': num_labels_gt,
                'obj_detected_075_percentage': pct}

    if num_labels_pred == 0 and num_labels_gt > 0:       # all false negatives (:139-150)
        return edge(0., 1., 0., 0.)
    if num_labels_pred > 0 and num_labels_gt == 0:       # all false positives (:151-162)
        return edge(0., 0., 1., 0.)
    if num_labels_pred == 0 and num_labels_gt == 0:      # correctly predicted nothing (:163-174)
        return edge(1., 1., 1., 1.)

    ig, ip = np.array(labels_gt), np.array(labels_pred)
    true_positives = tp_all[np.ix_(ig, ip)].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        prec = true_positives / area_pred[ip][None, :]
        rec = true_positives / area_gt[ig][:, None]
        F = np.where(prec + rec > 0, (2 * prec * rec) / (prec + rec), 0.0)
    boundary_stuff = np.stack([bprec_all[np.ix_(ig, ip)], brec_all[np.ix_(ig, ip)]], axis=2).astype(np.float64)
    F[np.isnan(F)] = 0
    assignments = _assignment(F.max() - F.copy())
    num_obj_detected = sum(1 for a in assignments if F[a] > obj_detect_threshold)
    idx = tuple(np.array(assignments).T)
    with np.errstate(divide='ignore', invalid='ignore'):
        precision = np.sum(true_positives[idx]) / np.float64(area_pred[1:].sum())     # prediction.clip(0,1) == 1
        recall = np.sum(true_positives[idx]) / np.float64(area_gt[1:].sum())
        F_measure = (2 * precision * recall) / (precision + recall)
        if np.isnan(F_measure):
            F_measure = 0
        boundary_precision = np.sum(boundary_stuff[idx][:, 0]) / np.float64(bden_p)
        boundary_recall = np.sum(boundary_stuff[idx][:, 1]) / np.float64(bden_g)
        boundary_F_measure = (2 * boundary_precision * boundary_recall) / (boundary_precision + boundary_recall)
        if np.isnan(boundary_F_measure):
            boundary_F_measure = 0
    return {'Objects F-measure': F_measure, 'Objects Precision': precision, 'Objects Recall': recall,
            'Boundary F-measure': boundary_F_measure, 'Boundary Precision': boundary_precision,
            'Boundary Recall': boundary_recall, 'obj_detected': num_labels_pred, 'obj_detected_075': num_obj_detected,
            'obj_gt': num_labels_gt, 'obj_detected_075_percentage': num_obj_detected / num_labels_gt}
