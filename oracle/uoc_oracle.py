"""CPU oracle for the UnseenObjectClustering inference hot path  --  TEST INFRASTRUCTURE ONLY.

A plain torch-CPU / numpy restatement of the reference algorithm (NVlabs/UnseenObjectClustering,
commit f5a00c7).  It is the checker the parity tests compare the CUDA path against, and the
`cpu_baseline` / `--impl reference` leg of bench.py.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors for this path, so
this file is pinned against the reference ITSELF, imported in the build container through
oracle/ref_harness.py: tests/test_oracle_vs_reference.py (skipped where /root/reference is
absent) and the fixtures under tests/golden/ written by oracle/make_golden.py.

Each function cites the reference file:line it follows (paths relative to the reference root).
The arithmetic deliberately uses the same library calls as the reference (torch.mm, torch.exp,
F.normalize, torch.argmax ...) so that its CPU timing is representative of the reference's CPU
path and its results are bit-identical to the reference on CPU.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EMBEDDING_ALPHA = 0.02      # lib/fcn/config.py:254  (cfg.TRAIN.EMBEDDING_ALPHA)
KAPPA = 20.0                # lib/fcn/test_dataset.py:51
MAX_ITERS = 10              # lib/fcn/test_dataset.py:56
NUM_SEEDS = 100             # lib/fcn/test_dataset.py:248 (stage 1) and :44 default (stage 2)
CROP_SIZE = 224             # lib/fcn/config.py:129 (cfg.TRAIN.SYN_CROP_SIZE)
BN_EPS = 1e-5               # torch.nn.BatchNorm2d default, lib/networks/resnet.py:50


# --------------------------------------------------------------------------------------------
# mean shift clustering  (lib/utils/mean_shift.py)
# --------------------------------------------------------------------------------------------

def cosine_kernel(Z, X, kappa):
    """lib/utils/mean_shift.py:25-26  W = exp(kappa * Z X^T)   ([m,d],[n,d] -> [m,n])."""
    return torch.exp(kappa * torch.mm(Z, X.t()))


def label_mode(values):
    """lib/utils/mean_shift.py:30-38  most frequent value, ties -> smallest value."""
    vals, counts = np.unique(values, return_counts=True)
    return int(vals[np.argmax(counts)])


def ball_kernel(Z, X, kappa, metric="cosine"):
    """lib/utils/mean_shift.py:11-27, both branches: euclidean W = exp(-kappa * ||z - x||^2) with the norm taken
    first and squared afterwards (:22-24, materialises [m, n, d]); cosine W = exp(kappa * Z X^T) (:26)."""
    if metric == "euclidean":
        distance = torch.norm(Z.unsqueeze(1) - X.unsqueeze(0), dim=2)
        return torch.exp(-kappa * torch.pow(distance, 2))
    return cosine_kernel(Z, X, kappa)


def _seed_distances(Zall, z, metric):
    """distance of every row of Zall to the single row z ([1,d]): mean_shift.py:58-62 / :159-162 / :180-184."""
    if metric == "euclidean":
        return torch.norm(Zall.unsqueeze(1) - z.unsqueeze(0), dim=2)[:, 0]
    return 0.5 * (1 - torch.mm(Zall, z.t()))[:, 0]


def label_seeds(Z, epsilon, metric="cosine"):
    """lib/utils/mean_shift.py:41-76.  Greedy, order dependent labelling of the
    converged seeds: every still-unlabelled seed i claims all seeds within cosine distance epsilon
    (fp32 `<=`); the claimed set takes the mode of its existing labels if any member is already
    labelled, else a fresh label; the WHOLE claimed set is overwritten."""
    m = Z.shape[0]
    labels = torch.full((m,), -1, dtype=torch.long)
    next_label = 0
    for i in range(m):
        if labels[i] != -1:
            continue
        dist = _seed_distances(Z, Z[i:i + 1], metric)
        comp = dist <= epsilon
        current = labels[comp]
        if torch.unique(current).shape[0] > 1:
            cur = current.numpy()
            lab = label_mode(cur[cur != -1])
        else:
            lab = next_label
            next_label += 1
        labels[comp] = lab
    return labels


def hill_climb(X, Z, kappa, max_iters, metric="cosine"):
    """lib/utils/mean_shift.py:79-109.  max_iters fixed-count mean-shift updates, no convergence test:
    cosine Z <- normalize_rows(exp(kappa Z X^T) X) (:107); euclidean Z <- (W X) / clamp(W.sum(1), min=1) (:101-105)."""
    for _ in range(max_iters):
        W = ball_kernel(Z, X, kappa, metric)
        if metric == "euclidean":
            Z = torch.mm(W, X) / torch.clamp(W.sum(dim=1).unsqueeze(1), min=1.0)
        else:
            Z = F.normalize(torch.mm(W, X), p=2, dim=1)
    return Z


def select_seeds(X, num_seeds, first_index, metric="cosine"):
    """lib/utils/mean_shift.py:128-189.  Farthest point sampling (euclidean: ||X - seed||, :159-160,:181-182).  `first_index`
    is the value the reference draws with np.random.randint(0, n) (:155).  Returns
    (seeds [m,d], selected_indices [m] int64)."""
    n, d = X.shape
    selected = -torch.ones(num_seeds, dtype=torch.long)
    seeds = torch.empty((num_seeds, d))
    distances = torch.empty((n, num_seeds))
    selected[0] = int(first_index)
    seeds[0] = X[int(first_index)]
    def column(seed):
        if metric == "euclidean":
            return torch.norm(X - seed.unsqueeze(0), dim=1)
        return 0.5 * (1 - torch.mm(X, seed.unsqueeze(1))[:, 0])

    distances[:, 0] = column(seeds[0])
    for i in range(1, num_seeds):
        nearest = torch.min(distances[:, :i], dim=1)[0]
        idx = torch.argmax(nearest)
        selected[i] = idx
        seeds[i] = X[idx]
        distances[:, i] = column(seeds[i])
    return seeds, selected


def select_seeds_init(X, num_seeds, init_seeds, num_init_seeds, metric="cosine"):
    """lib/utils/mean_shift.py:128-189 with init_seeds / num_init_seeds (:144-149, :164-169): the first num_init_seeds rows of
    init_seeds are seeds chosen before; their distance columns are computed first, sampling continues from them.  Like the
    reference, init_seeds is filled in place and returned; selected_indices stays -1 for the given rows (:140).
    Returns (seeds [m,d] (= init_seeds), selected_indices [m] int64)."""
    n, d = X.shape
    selected = -torch.ones(num_seeds, dtype=torch.long)
    seeds = init_seeds
    distances = torch.empty((n, num_seeds))

    def column(seed_row):                               # seed_row: [1, d]
        if metric == "euclidean":
            return torch.norm(X - seed_row, dim=1)
        return 0.5 * (1 - torch.mm(X, seed_row.t())[:, 0])

    for i in range(num_init_seeds):
        distances[:, i] = column(seeds[i:i + 1, :])
    for i in range(num_init_seeds, num_seeds):
        nearest = torch.min(distances[:, :i], dim=1)[0]
        idx = torch.argmax(nearest)
        selected[i] = idx
        seeds[i, :] = torch.index_select(X, 0, idx)[0, :]
        distances[:, i] = column(seeds[i:i + 1, :])
    return seeds, selected


def assign_and_relabel(X, Z, seed_labels, metric="cosine"):
    """lib/utils/mean_shift.py:206-227.  Nearest-seed assignment (argmin of 0.5(1 - x.z), first
    minimum), then swap label 0 with the most populated label.  Reproduces the histogram quirk:
    counts exist only for labels in range(len(unique(seed_labels)))."""
    if metric == "euclidean":                                   # mean_shift.py:207-209
        dist = torch.norm(X.unsqueeze(1) - Z.unsqueeze(0), dim=2)
    else:
        dist = 0.5 * (1 - torch.mm(X, Z.t()))
    closest = torch.argmin(dist, dim=1)
    labels = seed_labels[closest]
    num = len(torch.unique(seed_labels))
    count = torch.zeros(num, dtype=torch.long)
    for i in range(num):
        count[i] = (labels == i).sum()
    label_max = int(torch.argmax(count))
    if label_max != 0:
        idx0 = labels == 0
        idx1 = labels == label_max
        labels[idx0] = label_max
        labels[idx1] = 0
    return labels


def mean_shift_smart_init(X, kappa=KAPPA, num_seeds=NUM_SEEDS, max_iters=MAX_ITERS, first_index=None,
                          return_all=False, metric="cosine"):
    """lib/utils/mean_shift.py:192-229.  X: [n,d] unit rows (any strides).  `first_index` None ->
    draw it from numpy's global RNG exactly like the reference (:155)."""
    n = X.shape[0]
    if first_index is None:
        first_index = np.random.randint(0, n)
    seeds, selected = select_seeds(X, num_seeds, first_index, metric)
    Z = hill_climb(X, seeds, kappa, max_iters, metric)
    seed_labels = label_seeds(Z, 2 * EMBEDDING_ALPHA, metric)  # mean_shift.py:123
    labels = assign_and_relabel(X, Z, seed_labels, metric)
    if return_all:
        return labels, selected, seeds, Z, seed_labels
    return labels, selected


def clustering_features(features, num_seeds=NUM_SEEDS, first_indices=None, metric="cosine"):
    """lib/fcn/test_dataset.py:44-59.  features [N,C,H,W] -> (labels float32 [N,H,W] CPU,
    list of N int64 [num_seeds] tensors).  metric = cfg.TRAIN.EMBEDDING_METRIC (:45)."""
    N, C, H, W = features.shape
    out = torch.zeros((N, H, W))
    picked = []
    for j in range(N):
        X = features[j].reshape(C, -1).t()
        fi = None if first_indices is None else int(first_indices[j])
        labels, sel = mean_shift_smart_init(X, KAPPA, num_seeds, MAX_ITERS, fi, metric=metric)
        out[j] = labels.view(H, W)
        picked.append(sel)
    return out, picked


# --------------------------------------------------------------------------------------------
# two-stage plumbing  (lib/fcn/test_dataset.py, lib/utils/mask.py)
# --------------------------------------------------------------------------------------------

def filter_labels_depth(labels, depth, threshold):
    """lib/fcn/test_dataset.py:183-198.  Zero every non-background id whose fraction of pixels
    with Z > 0 (depth channel 2) is below `threshold`."""
    out = labels.clone()
    for i in range(labels.shape[0]):
        ids = torch.unique(labels[i])
        for mid in ids:
            if mid == 0:
                continue
            sel = labels[i] == mid
            frac = (depth[i, 2][sel] > 0).sum().float() / sel.sum().float()
            if frac < threshold:
                out[i][sel] = 0
    return out


def tight_box(mask):
    """lib/utils/mask.py:180-187  -> (x_min, y_min, x_max, y_max) of the non-zeros."""
    nz = torch.nonzero(mask)
    return int(nz[:, 1].min()), int(nz[:, 0].min()), int(nz[:, 1].max()), int(nz[:, 0].max())


def _round_half_even(v):
    return int(torch.round(torch.tensor(float(v))).item())


def crop_rois(rgb, initial_masks, depth, crop_size=CROP_SIZE):
    """lib/fcn/test_dataset.py:62-112.  Only batch item 0 is cropped (:68,97,100).  ROI = tight
    box padded by round(0.25 * extent) (torch.round: half to even, extent without +1), clamped;
    rgb/depth resized bilinear align_corners=True, mask nearest (legacy floor)."""
    N, H, W = initial_masks.shape
    ids = torch.unique(initial_masks[0])
    ids = ids[ids != 0] if (len(ids) and ids[0] == 0) else ids
    K = ids.shape[0]
    rgb_crops = torch.zeros((K, 3, crop_size, crop_size))
    mask_crops = torch.zeros((K, crop_size, crop_size))
    rois = torch.zeros((K, 4))
    depth_crops = torch.zeros((K, 3, crop_size, crop_size)) if depth is not None else None
    size = (crop_size, crop_size)
    for k, mid in enumerate(ids):
        mask = (initial_masks[0] == mid).float()
        x0, y0, x1, y1 = tight_box(mask)
        xp = _round_half_even((x1 - x0) * 0.25)
        yp = _round_half_even((y1 - y0) * 0.25)
        x0 = max(x0 - xp, 0); x1 = min(x1 + xp, W - 1)
        y0 = max(y0 - yp, 0); y1 = min(y1 + yp, H - 1)
        rois[k] = torch.tensor([x0, y0, x1, y1], dtype=torch.float32)
        rgb_crops[k] = F.interpolate(rgb[0:1, :, y0:y1 + 1, x0:x1 + 1], size=size, mode="bilinear",
                                     align_corners=True)[0]
        mask_crops[k] = F.interpolate(mask[None, None, y0:y1 + 1, x0:x1 + 1], size=size, mode="nearest")[0, 0]
        if depth is not None:
            depth_crops[k] = F.interpolate(depth[0:1, :, y0:y1 + 1, x0:x1 + 1], size=size, mode="bilinear",
                                           align_corners=True)[0]
    return rgb_crops, mask_crops, rois, depth_crops


def match_label_crop(initial_masks, labels_crop, mask_crops, rois, depth_crops):
    """lib/fcn/test_dataset.py:116-179.  (i) drop (-1) crop clusters overlapping the stage-1 mask by
    < 50 % of their area; (ii) order crops far -> near by mean Z of kept pixels with Z > 0 (ROI area
    when no depth), descending; (iii) renumber kept clusters 1,2,3.. in that order, nearest-resize
    to the ROI and paste the non-zeros, nearer crops overwriting.  Mutates labels_crop like the
    reference (:125)."""
    K = labels_crop.shape[0]
    for i in range(K):
        for mid in torch.unique(labels_crop[i]):
            sel = labels_crop[i] == mid
            frac = (sel.float() * mask_crops[i]).sum() / sel.float().sum()
            if frac < 0.5:
                labels_crop[i][sel] = -1
    keyed = []
    for i in range(K):
        if depth_crops is not None:
            kept = labels_crop[i] > -1
            zs = depth_crops[i, 2][kept] if kept.sum() > 0 else depth_crops[i, 2]
            keyed.append((i, torch.mean(zs[zs > 0])))
        else:
            keyed.append((i, (rois[i, 3] - rois[i, 1] + 1) * (rois[i, 2] - rois[i, 0] + 1)))
    order = [i for i, _ in sorted(keyed, key=lambda t: t[1], reverse=True)]
    refined = torch.zeros_like(initial_masks).float()
    count = 0
    for i in order:
        ids = torch.unique(labels_crop[i])
        ids = ids[1:] if ids[0] == -1 else ids
        renum = torch.zeros_like(labels_crop[i])
        for mid in ids:
            count += 1
            renum[labels_crop[i] == mid] = count
        x0, y0, x1, y1 = (int(rois[i, j].item()) for j in range(4))
        back = F.interpolate(renum[None, None].float(), size=(y1 - y0 + 1, x1 - x0 + 1), mode="nearest")[0, 0]
        hh, ww = torch.nonzero(back).t()
        refined[0, y0:y1 + 1, x0:x1 + 1][hh, ww] = back[hh, ww]
    return refined, labels_crop


def test_sample(image, depth, network, network_crop, first_indices=None, first_indices_crop=None):
    """lib/fcn/test_dataset.py:232-267 on CPU tensors.  `network(image, None, depth)` must return
    unit-norm [N,C,H,W] features.  Returns (out_label, out_label_refined | None)."""
    features = network(image, None, depth).detach()
    out_label, _ = clustering_features(features, NUM_SEEDS, first_indices)
    if depth is not None:
        out_label = filter_labels_depth(out_label, depth, 0.8)
    refined = None
    if network_crop is not None:
        rgb_c, mask_c, rois, depth_c = crop_rois(image, out_label.clone(), depth)
        if rgb_c.shape[0] > 0:
            feats_c = network_crop(rgb_c, mask_c, depth_c).detach()
            labels_c, _ = clustering_features(feats_c, NUM_SEEDS, first_indices_crop)
            refined, _ = match_label_crop(out_label, labels_c, mask_c, rois, depth_c)
    return out_label, refined


# --------------------------------------------------------------------------------------------
# backbone: ResNet34-8s two-branch RGB-D add-fusion  (lib/networks/{SEG,resnet_dilated,resnet}.py)
# --------------------------------------------------------------------------------------------

# (planes, blocks, first-block stride, dilation) after the output_stride=8 stride->dilation
# conversion of lib/networks/resnet.py:188-234 (layer3/4: stride 1, dilation 2 / 4).
RESNET34_8S_LAYERS = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2), (512, 3, 1, 4))


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def resnet34_8s_trunk(x, sd, prefix):
    """lib/networks/resnet.py:236-270 with fully_conv / remove_avg_pool_layer / output_stride=8
    (lib/networks/resnet_dilated.py:296-303): stem 7x7 s2 -> BN -> ReLU -> maxpool 3x3 s2 ->
    4 stages of BasicBlock (resnet.py:57-73) -> 1x1 conv `fc` with bias.  Eval-mode BN."""
    p = prefix
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, p + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, (planes, blocks, stride, dil) in enumerate(RESNET34_8S_LAYERS, start=1):
        for b in range(blocks):
            q = "%slayer%d.%d." % (p, li, b)
            s = stride if b == 0 else 1
            out = F.conv2d(x, sd[q + "conv1.weight"], None, stride=s, padding=dil, dilation=dil)
            out = F.relu(_bn(out, sd, q + "bn1"))
            out = F.conv2d(out, sd[q + "conv2.weight"], None, stride=1, padding=dil, dilation=dil)
            out = _bn(out, sd, q + "bn2")
            if (q + "downsample.0.weight") in sd:
                res = _bn(F.conv2d(x, sd[q + "downsample.0.weight"], None, stride=s), sd, q + "downsample.1")
            else:
                res = x
            x = F.relu(out + res)
    return F.conv2d(x, sd[p + "fc.weight"], sd[p + "fc.bias"])


def segnet_rgbd_add_forward(sd, img, depth):
    """lib/networks/SEG.py:88-119 (INPUT='RGBD', FUSION_TYPE='add', eval): each branch = trunk ->
    bilinear upsample to the input size with align_corners=True (resnet_dilated.py:325), add
    (:108), F.normalize over channels (:114).  `sd` is a reference-format state_dict."""
    size = img.shape[2:]
    a = F.interpolate(resnet34_8s_trunk(img, sd, "fcn.resnet34_8s."), size=size, mode="bilinear", align_corners=True)
    b = F.interpolate(resnet34_8s_trunk(depth, sd, "fcn_depth.resnet34_8s."), size=size, mode="bilinear",
                      align_corners=True)
    return F.normalize(a + b, p=2, dim=1)


def normalize_state_dict(data):
    """Key handling of lib/networks/SEG.py:130-159 and tools/test_net.py:111-112: unwrap
    {'model': sd}, strip a leading 'module.'."""
    if isinstance(data, dict) and "model" in data and not torch.is_tensor(data["model"]):
        data = data["model"]
    out = {}
    for k, v in data.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def segnet_forward(sd, img, depth, input_type="RGBD", fusion_type="add", normalize=True):
    """lib/networks/SEG.py:88-119 in eval mode for every input / fusion variant of the ResNet34-8s network:
    DEPTH -> fcn(depth) (:97-98); COLOR -> fcn(img) (:99-100); RGBD early -> fcn(cat(img, depth)) (:101-103);
    RGBD add / cat -> fcn(img) (+ | cat) fcn_depth(depth) (:105-110); F.normalize iff EMBEDDING_NORMALIZATION (:113-114)."""
    def branch(x, prefix):
        return F.interpolate(resnet34_8s_trunk(x, sd, prefix), size=x.shape[2:], mode="bilinear", align_corners=True)

    if input_type == "DEPTH":
        f = branch(depth, "fcn.resnet34_8s.")
    elif input_type == "COLOR":
        f = branch(img, "fcn.resnet34_8s.")
    elif input_type == "RGBD" and fusion_type == "early":
        f = branch(torch.cat((img, depth), 1), "fcn.resnet34_8s.")
    else:
        a = branch(img, "fcn.resnet34_8s.")
        b = branch(depth, "fcn_depth.resnet34_8s.")
        f = a + b if fusion_type == "add" else torch.cat((a, b), 1)
    return F.normalize(f, p=2, dim=1) if normalize else f


class OracleSegNet:
    """Callable with the reference's network signature net(img, label, depth) -> features."""

    def __init__(self, state_dict, input_type="RGBD", fusion_type="add", normalize=True):
        self.sd = {k: v.detach().float() for k, v in normalize_state_dict(state_dict).items()}
        self.variant = (input_type, fusion_type, normalize)

    def __call__(self, img, label=None, depth=None):
        with torch.no_grad():
            if self.variant == ("RGBD", "add", True):
                return segnet_rgbd_add_forward(self.sd, img, depth)
            return segnet_forward(self.sd, img, depth, *self.variant)


# --------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# --------------------------------------------------------------------------------------------

def synthetic_clustered_features(H, W, d=64, num_objects=6, noise=0.05, seed=0):
    """cfg2 generator: unit-norm embedding field [1,d,H,W] with num_objects axis-aligned rectangles
    on a background; x = normalize(centre[label] + noise * randn).  Returns (features, gt [H,W])."""
    g = torch.Generator().manual_seed(seed)
    centres = F.normalize(torch.randn(num_objects + 1, d, generator=g), dim=1)
    gt = torch.zeros(H, W, dtype=torch.long)
    for k in range(1, num_objects + 1):
        h = int(torch.randint(H // 8, H // 3, (1,), generator=g))
        w = int(torch.randint(W // 8, W // 3, (1,), generator=g))
        y = int(torch.randint(0, H - h, (1,), generator=g))
        x = int(torch.randint(0, W - w, (1,), generator=g))
        gt[y:y + h, x:x + w] = k
    X = centres[gt.view(-1)] + noise * torch.randn(H * W, d, generator=g)
    X = F.normalize(X, dim=1)
    feats = X.t().contiguous().view(1, d, H, W)
    return feats, gt


def exact_clustered_field(H, W, d=64, num_objects=6, noise=0.05, seed=0, chunk_rows=32768):
    """Platform-independent cfg2 / cfg5 generator for the FULL-SIZE fixtures (tests/golden/full_*.npz store only the
    reference's outputs; the 79 - 354 MB inputs are regenerated from the seed on the box that runs the test, so the
    generator must give the same bits everywhere): MT19937 integers only; noise = Irwin-Hall sum of four 16-bit uniforms
    (an integer); rows are quantised to 2^-20 and normalised by an EXACT integer sum of squares, so every float
    operation is a single correctly rounded IEEE op (no vectorised reductions whose order depends on the CPU).
    Returns (X_planar float32 [d, H*W] (row k = channel k, the reference's memory layout), gt int64 [H, W])."""
    rs = np.random.RandomState(int(seed))
    ci = rs.randint(-1000, 1001, size=(num_objects + 1, d)).astype(np.int64)
    c = ci / np.sqrt((ci * ci).sum(1).astype(np.float64))[:, None]
    gt = np.zeros((H, W), dtype=np.int64)
    for k in range(1, num_objects + 1):
        h = int(rs.randint(H // 8, H // 3)); w = int(rs.randint(W // 8, W // 3))
        y = int(rs.randint(0, H - h)); x = int(rs.randint(0, W - w))
        gt[y:y + h, x:x + w] = k
    n = H * W
    flat = gt.reshape(-1)
    inv_sigma = 1.0 / math.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0)
    out = np.empty((d, n), dtype=np.float32)
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        u = rs.randint(0, 65536, size=(r1 - r0, d, 4)).astype(np.int64).sum(-1)
        z = (u - 131070).astype(np.float64) * inv_sigma
        x = c[flat[r0:r1]] + noise * z
        xi = np.rint(x * 1048576.0).astype(np.int64)
        ss = (xi * xi).sum(1)
        out[:, r0:r1] = (xi / np.sqrt(ss.astype(np.float64))[:, None]).astype(np.float32).T
    return out, gt


def field_crc32(X_planar):
    import zlib
    return zlib.crc32(np.ascontiguousarray(X_planar).tobytes()) & 0xFFFFFFFF


def synthetic_rgbd_frame(H=480, W=640, seed=0):
    """cfg1 generator: image ~ U(-0.5, 0.6); XYZ from Z ~ U(0.3, 1.5) m with the demo intrinsics
    (data/demo/camera_params.json: fx 612.937 fy 613.173 cx 322.549 cy 248.158, scaled to HxW)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, H, W, generator=g) * 1.1 - 0.5
    z = torch.rand(1, 1, H, W, generator=g) * 1.2 + 0.3
    sx, sy = W / 640.0, H / 480.0
    fx, fy, cx, cy = 612.937 * sx, 613.173 * sy, 322.549 * sx, 248.158 * sy
    u = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    v = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xyz = torch.cat([(u - cx) / fx * z, (v - cy) / fy * z, z], dim=1)
    return img, xyz


def structured_rgbd_frame(H=480, W=640, num_objects=6, seed=0):
    """A frame with STRUCTURE for the label flip-rate test (SURVEY section 7 "report the flip rate"): num_objects
    rectangles on a background, each region with its own constant colour and its own constant 3-vector in the depth
    tensor (piece-wise constant network inputs; not a physical XYZ map -- spatial ramps would give a continuum of
    embeddings).  Returns (img [1,3,H,W], xyz [1,3,H,W], gt [H,W])."""
    g = torch.Generator().manual_seed(int(seed))
    gt = torch.zeros(H, W, dtype=torch.long)
    for k in range(1, num_objects + 1):
        h = int(torch.randint(H // 6, H // 3, (1,), generator=g))
        w = int(torch.randint(W // 6, W // 3, (1,), generator=g))
        y = int(torch.randint(0, H - h, (1,), generator=g))
        x = int(torch.randint(0, W - w, (1,), generator=g))
        gt[y:y + h, x:x + w] = k
    col = torch.rand(num_objects + 1, 3, generator=g) * 1.1 - 0.5
    pc = torch.rand(num_objects + 1, 3, generator=g) * 1.0 + 0.2
    img = col[gt].permute(2, 0, 1)[None].contiguous()
    xyz = pc[gt].permute(2, 0, 1)[None].contiguous()
    return img, xyz, gt


def calibrated_state_dict_(sd, img, xyz, residual_gain=0.05):
    """Turn a random-init reference-format state_dict into one whose embeddings of (img, xyz) are NOT collapsed (random
    init gives one cluster): the residual branches are scaled down (bn2.weight *= residual_gain, so the
    identity / down-sample path carries the regions' piece-wise constant signal), every BatchNorm gets the batch
    statistics of this very input as running statistics (layer by layer, as a training-mode pass would accumulate them),
    and the fc bias centres the trunk output.  In place; returns sd.  Test infrastructure (fp32 torch on the CPU)."""
    for k in list(sd.keys()):
        if k.endswith("bn2.weight"):
            sd[k] = sd[k] * residual_gain

    def bn(x, p):
        m = x.mean((0, 2, 3))
        v = x.var((0, 2, 3), unbiased=False)
        sd[p + ".running_mean"] = m.clone()
        sd[p + ".running_var"] = v.clone()
        return F.batch_norm(x, m, v, sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)

    with torch.no_grad():
        for x, p in ((img, "fcn.resnet34_8s."), (xyz, "fcn_depth.resnet34_8s.")):
            x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
            x = F.max_pool2d(F.relu(bn(x, p + "bn1")), kernel_size=3, stride=2, padding=1)
            for li, (planes, blocks, stride, dil) in enumerate(RESNET34_8S_LAYERS, start=1):
                for b in range(blocks):
                    q = "%slayer%d.%d." % (p, li, b)
                    s = stride if b == 0 else 1
                    out = F.relu(bn(F.conv2d(x, sd[q + "conv1.weight"], None, stride=s, padding=dil, dilation=dil), q + "bn1"))
                    out = bn(F.conv2d(out, sd[q + "conv2.weight"], None, stride=1, padding=dil, dilation=dil), q + "bn2")
                    if (q + "downsample.0.weight") in sd:
                        res = bn(F.conv2d(x, sd[q + "downsample.0.weight"], None, stride=s), q + "downsample.1")
                    else:
                        res = x
                    x = F.relu(out + res)
            y = F.conv2d(x, sd[p + "fc.weight"], None)
            sd[p + "fc.bias"] = -y.mean((0, 2, 3))
    return sd


def best_label_agreement(a, b):
    """Fraction of pixels on which label maps a and b agree under the best one-to-one relabelling of b that keeps
    label 0 fixed (Hungarian assignment on the contingency table of the non-zero ids)."""
    from scipy.optimize import linear_sum_assignment
    a = np.asarray(a).astype(np.int64).ravel()
    b = np.asarray(b).astype(np.int64).ravel()
    ka, kb = int(a.max()) + 1, int(b.max()) + 1
    table = np.zeros((ka, kb), dtype=np.int64)
    np.add.at(table, (a, b), 1)
    agree = int(table[0, 0])
    if ka > 1 and kb > 1:
        sub = table[1:, 1:]
        r, c = linear_sum_assignment(-sub)
        agree += int(sub[r, c].sum())
    return agree / float(a.size)


def randomise_bn_(sd, seed):
    """Give every BatchNorm of a reference-format state_dict non-trivial statistics / affine terms
    (and the fc a bias) so that BN folding is exercised.  In place; returns sd."""
    g = torch.Generator().manual_seed(int(seed))
    for k in list(sd.keys()):
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
        elif k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
        elif k.endswith("bn1.weight") or k.endswith("bn2.weight") or k.endswith("downsample.1.weight"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.4 + 0.8
        elif k.endswith("bn1.bias") or k.endswith("bn2.bias") or k.endswith("downsample.1.bias") or k.endswith("fc.bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    return sd


def labels_equal_up_to_permutation(a, b):
    """True iff integer label maps a, b are identical up to a relabelling that fixes label 0
    (0 = background / largest cluster is special downstream)."""
    a = np.asarray(a).astype(np.int64).ravel()
    b = np.asarray(b).astype(np.int64).ravel()
    if a.shape != b.shape:
        return False
    if not np.array_equal(a == 0, b == 0):
        return False
    fwd, bwd = {}, {}
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    for x, y in pairs:
        if fwd.setdefault(int(x), int(y)) != int(y) or bwd.setdefault(int(y), int(x)) != int(x):
            return False
    return True


# ---------------------------------------------------------------------------------------------
# input preparation (the step before the path)
# ---------------------------------------------------------------------------------------------
PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])        # fcn/config.py:376


def compute_xyz(depth_img, fx, fy, px, py, height, width):
    """tools/test_images.py:96-102 (utils/mask.py:41-46 for the index grid): [H,W] fp32 metres -> [H,W,3] fp32."""
    indices = np.indices((height, width), dtype=np.float32).transpose(1, 2, 0)
    z_e = depth_img
    x_e = (indices[..., 1] - px) * z_e / fx
    y_e = (indices[..., 0] - py) * z_e / fy
    return np.stack([x_e, y_e, z_e], axis=-1)


def read_sample_arrays(im_bgr_u8, depth_u16, camera_params, pixel_means=PIXEL_MEANS):
    """tools/test_images.py:105-135 after the two cv2.imread calls: (image_color [1,3,H,W], depth [1,3,H,W]) fp32 CPU."""
    depth = depth_u16.astype(np.float32) / 1000.0
    h, w = depth.shape
    xyz = compute_xyz(depth, camera_params['fx'], camera_params['fy'], camera_params['x_offset'],
                      camera_params['y_offset'], h, w)
    im_tensor = torch.from_numpy(im_bgr_u8) / 255.0
    im_tensor -= torch.tensor(pixel_means / 255.0).float()
    return im_tensor.permute(2, 0, 1).unsqueeze(0), torch.from_numpy(xyz).permute(2, 0, 1).unsqueeze(0)



# --------------------------------------------------------------------------------------------
# evaluation tail  (lib/utils/evaluation.py)
# --------------------------------------------------------------------------------------------

def disk(radius):
    """skimage.morphology.disk as used at lib/utils/evaluation.py:94-98 (skimage is not installed here):
    a (2r+1)x(2r+1) uint8 footprint of the pixels with x^2 + y^2 <= r^2."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.uint8)


def seg2bmap(seg):
    """lib/utils/evaluation.py:15-70 at full resolution: one-pixel boundary map of a binary mask."""
    seg = seg.astype(bool)
    e = np.zeros_like(seg)
    s = np.zeros_like(seg)
    se = np.zeros_like(seg)
    e[:, :-1] = seg[:, 1:]
    s[:-1, :] = seg[1:, :]
    se[:-1, :-1] = seg[1:, 1:]
    b = seg ^ e | seg ^ s | seg ^ se
    b[-1, :] = seg[-1, :] ^ e[-1, :]
    b[:, -1] = seg[:, -1] ^ s[:, -1]
    b[-1, -1] = 0
    return b


def boundary_overlap(predicted_mask, gt_mask, bound_th=0.003):
    """lib/utils/evaluation.py:73-106: (precision true positives, recall true positives) of the dilated boundaries."""
    import cv2
    bound_pix = bound_th if bound_th >= 1 else np.ceil(bound_th * np.linalg.norm(predicted_mask.shape))
    fg_boundary = seg2bmap(predicted_mask)
    gt_boundary = seg2bmap(gt_mask)
    gt_dil = cv2.dilate(gt_boundary.astype(np.uint8), disk(bound_pix), iterations=1)
    fg_dil = cv2.dilate(fg_boundary.astype(np.uint8), disk(bound_pix), iterations=1)
    return np.sum(np.logical_and(fg_boundary, gt_dil)), np.sum(np.logical_and(gt_boundary, fg_dil))


def munkres_assignment(cost):
    """lib/utils/munkres.py:320-372 (Munkres.compute) restated with plain loops: the Kuhn-Munkres steps in the reference's
    scan orders (zero padding to a square :301-316, row minima only :385-399, row-major greedy stars :401-418, the zero to
    prime = LAST uncovered zero of the first uncovered row that has one :536-560, update = add to covered rows then subtract
    from uncovered columns :510-524, first star in column / first prime in row :562-599), so that ties are broken alike.
    Returns the (row, col) pairs inside the original matrix, row-major."""
    cost = np.asarray(cost, dtype=np.float64)
    R, Cn = cost.shape
    n = max(R, Cn)
    C = [[float(cost[i, j]) if (i < R and j < Cn) else 0.0 for j in range(n)] for i in range(n)]
    for i in range(n):
        mv = min(C[i])
        for j in range(n):
            C[i][j] -= mv
    mark = [[0] * n for _ in range(n)]               # 1 star, 2 prime
    rc, cc = [False] * n, [False] * n
    for i in range(n):
        for j in range(n):
            if C[i][j] == 0 and not rc[i] and not cc[j]:
                mark[i][j] = 1
                rc[i] = cc[j] = True
    rc, cc = [False] * n, [False] * n
    while True:
        count = 0
        for i in range(n):
            for j in range(n):
                if mark[i][j] == 1:
                    cc[j] = True
                    count += 1
        if count >= n:
            break
        while True:
            zr = zc = -1
            for i in range(n):
                found = False
                for j in range(n):
                    if C[i][j] == 0 and not rc[i] and not cc[j]:
                        zr, zc, found = i, j, True    # keeps scanning the row: the last zero wins
                if found:
                    break
            if zr < 0:
                mv = min(C[i][j] for i in range(n) for j in range(n) if not rc[i] and not cc[j])
                for i in range(n):
                    for j in range(n):
                        if rc[i]:
                            C[i][j] += mv
                        if not cc[j]:
                            C[i][j] -= mv
                continue
            mark[zr][zc] = 2
            sc = next((j for j in range(n) if mark[zr][j] == 1), -1)
            if sc < 0:
                break
            rc[zr] = True
            cc[sc] = False
        path = [(zr, zc)]
        while True:
            r = next((i for i in range(n) if mark[i][path[-1][1]] == 1), -1)
            if r < 0:
                break
            path.append((r, path[-1][1]))
            path.append((r, next(j for j in range(n) if mark[r][j] == 2)))
        for (i, j) in path:
            mark[i][j] = 0 if mark[i][j] == 1 else 1
        rc, cc = [False] * n, [False] * n
        for i in range(n):
            for j in range(n):
                if mark[i][j] == 2:
                    mark[i][j] = 0
    return [(i, j) for i in range(R) for j in range(Cn) if mark[i][j] == 1]


def multilabel_metrics(prediction, gt, obj_detect_threshold=0.75):
    """lib/utils/evaluation.py:109-257, statement by statement; the Hungarian step (utils.munkres, :221-223) through
    munkres_assignment above (the reference's procedure and tie-breaking)."""
    labels_gt = np.unique(gt)
    labels_gt = labels_gt[~np.isin(labels_gt, [0])]
    labels_pred = np.unique(prediction)
    labels_pred = labels_pred[~np.isin(labels_pred, [0])]
    ng, npred = labels_gt.shape[0], labels_pred.shape[0]

    def edge(f, p, r, pct):
        return {'Objects F-measure': f, 'Objects Precision': p, 'Objects Recall': r, 'Boundary F-measure': f,
                'Boundary Precision': p, 'Boundary Recall': r, 'obj_detected': npred, 'obj_detected_075': 0.,
                'obj_gt': ng, 'obj_detected_075_percentage': pct}

    if npred == 0 and ng > 0:
        return edge(0., 1., 0., 0.)
    if npred > 0 and ng == 0:
        return edge(0., 0., 1., 0.)
    if npred == 0 and ng == 0:
        return edge(1., 1., 1., 1.)
    F = np.zeros((ng, npred))
    true_positives = np.zeros((ng, npred))
    boundary_stuff = np.zeros((ng, npred, 2))
    for i, gt_i in enumerate(labels_gt):
        gt_i_mask = (gt == gt_i)
        for j, pred_j in enumerate(labels_pred):
            pred_j_mask = (prediction == pred_j)
            tp = np.int64(np.count_nonzero(np.logical_and(pred_j_mask, gt_i_mask)))
            true_positives[i, j] = tp
            prec = tp / np.count_nonzero(pred_j_mask)
            rec = tp / np.count_nonzero(gt_i_mask)
            if prec + rec > 0:
                F[i, j] = (2 * prec * rec) / (prec + rec)
            boundary_stuff[i, j] = boundary_overlap(pred_j_mask, gt_i_mask)
    boundary_prec_denom = sum(float(np.sum(seg2bmap(prediction == pred_j))) for pred_j in labels_pred)
    boundary_rec_denom = sum(float(np.sum(seg2bmap(gt == gt_i))) for gt_i in labels_gt)
    F[np.isnan(F)] = 0
    assignments = munkres_assignment(F.max() - F.copy())
    num_obj_detected = sum(1 for a in assignments if F[a] > obj_detect_threshold)
    idx = tuple(np.array(assignments).T)
    with np.errstate(divide='ignore', invalid='ignore'):
        precision = np.sum(true_positives[idx]) / np.sum(prediction.clip(0, 1) == 1)
        recall = np.sum(true_positives[idx]) / np.sum(gt.clip(0, 1) == 1)
        F_measure = (2 * precision * recall) / (precision + recall)
        if np.isnan(F_measure):
            F_measure = 0
        boundary_precision = np.sum(boundary_stuff[idx][:, 0]) / boundary_prec_denom
        boundary_recall = np.sum(boundary_stuff[idx][:, 1]) / boundary_rec_denom
        boundary_F_measure = (2 * boundary_precision * boundary_recall) / (boundary_precision + boundary_recall)
        if np.isnan(boundary_F_measure):
            boundary_F_measure = 0
    return {'Objects F-measure': F_measure, 'Objects Precision': precision, 'Objects Recall': recall,
            'Boundary F-measure': boundary_F_measure, 'Boundary Precision': boundary_precision,
            'Boundary Recall': boundary_recall, 'obj_detected': npred, 'obj_detected_075': num_obj_detected,
            'obj_gt': ng, 'obj_detected_075_percentage': num_obj_detected / ng}


def synthetic_prediction(gt, seed):
    """A plausible prediction for the label map `gt` ([H,W] ints): masks shifted by a few pixels, one object split in two,
    one dropped, one false positive, ids permuted."""
    rng = np.random.RandomState(seed)
    gt = np.asarray(gt)
    H, W = gt.shape
    pred = np.zeros_like(gt)
    ids = [l for l in np.unique(gt) if l != 0]
    nxt = 1
    for k, l in enumerate(ids):
        if k == 1 and len(ids) > 2 and seed % 2 == 0:
            continue                                   # a missed object
        m = np.roll(np.roll(gt == l, rng.randint(-3, 4), axis=0), rng.randint(-3, 4), axis=1)
        if k == 0 and seed % 3 != 2:                   # an over-segmented object
            ys = np.nonzero(m.any(1))[0]
            mid = (ys[0] + ys[-1]) // 2
            top = m.copy(); top[mid:] = False
            pred[top] = nxt; nxt += 1
            m = m & ~top
        pred[m] = nxt; nxt += 1
    y0, x0 = rng.randint(0, H - 8), rng.randint(0, W - 8)
    pred[y0:y0 + 6, x0:x0 + 7] = nxt                    # a false positive
    perm = rng.permutation(np.arange(1, nxt + 1)) + 3   # arbitrary ids, gaps included
    lut = np.zeros(nxt + 1, dtype=gt.dtype); lut[1:] = perm
    return lut[pred]
