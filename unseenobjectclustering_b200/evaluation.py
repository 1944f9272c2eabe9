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
    utils.munkres.Munkres().compute returns (lib/utils/evaluation.py:220-221, lib/utils/munkres.py:320-372) INCLUDING its
    choice among equally good assignments: the Kuhn-Munkres procedure with the reference's scan orders and float64
    arithmetic, restated on numpy state arrays.  What decides ties (lib/utils/munkres.py):
      * the smaller side is zero-padded to a square (:301-316); only ROW minima are subtracted (:385-399);
      * initial stars: row-major greedy over uncovered zeros (:401-418);
      * the zero to prime is taken in the FIRST uncovered row that has an uncovered zero, and in that row it is the LAST
        such zero -- the scan does not stop at the first hit (:536-560);
      * the matrix update adds the smallest uncovered value to covered rows, THEN subtracts it from uncovered columns,
        element by element (:510-524): the same two float64 operations in the same order here, so the exact == 0 tests see
        the same bits;
      * augmenting path: first star in the column, first prime in the row (:474-508, :562-599).
    Pinned against the reference's own class on thousands of tie-heavy matrices (tests/test_oracle.py) and against a
    fixture it wrote (tests/golden/munkres.npz)."""
    cost = np.asarray(cost, dtype=np.float64)
    rows, cols = cost.shape
    n = max(rows, cols)
    C = np.zeros((n, n), dtype=np.float64)
    C[:rows, :cols] = cost
    C -= C.min(axis=1, keepdims=True)
    star = np.zeros((n, n), dtype=bool)
    prime = np.zeros((n, n), dtype=bool)
    row_cov = np.zeros(n, dtype=bool)
    col_cov = np.zeros(n, dtype=bool)
    for i in range(n):                                   # initial stars
        free = np.flatnonzero((C[i] == 0) & ~col_cov)
        if free.size:
            star[i, free[0]] = True
            col_cov[free[0]] = True
    col_cov[:] = False
    while True:
        col_cov |= star.any(axis=0)
        if int(star.sum()) >= n:
            break
        z_row = z_col = -1
        while True:                                      # prime zeros until one has no star in its row
            open_zero = (C == 0) & ~row_cov[:, None] & ~col_cov[None, :]
            hit_rows = np.flatnonzero(open_zero.any(axis=1))
            if hit_rows.size == 0:                       # no uncovered zero: shift the matrix by the smallest uncovered value
                m = C[np.ix_(~row_cov, ~col_cov)].min()
                C[row_cov, :] += m
                C[:, ~col_cov] -= m
                continue
            i = int(hit_rows[0])
            j = int(np.flatnonzero(open_zero[i])[-1])
            prime[i, j] = True
            starred = np.flatnonzero(star[i])
            if starred.size:
                row_cov[i] = True
                col_cov[starred[0]] = False
            else:
                z_row, z_col = i, j
                break
        path = [(z_row, z_col)]                          # alternate: star in the column, prime in that star's row
        while True:
            above = np.flatnonzero(star[:, path[-1][1]])
            if above.size == 0:
                break
            r = int(above[0])
            path.append((r, path[-1][1]))
            path.append((r, int(np.flatnonzero(prime[r])[0])))
        for r, c in path:
            star[r, c] = not star[r, c]
        row_cov[:] = False
        col_cov[:] = False
        prime[:] = False
    return [(i, j) for i in range(rows) for j in range(cols) if star[i, j]]


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
