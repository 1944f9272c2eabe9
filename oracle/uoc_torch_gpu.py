"""The reference's own PyTorch path on a CUDA device  --  BASELINE / TEST INFRASTRUCTURE ONLY.

BASELINE.json's target is stated against "the reference's single-GPU PyTorch frames/sec on the same box".  The
reference tree does not exist on the GPU box, so this file restates, op for op, what the reference executes when its
tensors live on a GPU (cuDNN convolutions, cuBLAS `mm`, eager element-wise kernels, the same host round trips):

  lib/networks/SEG.py:88-119 + resnet_dilated.py:315-327 + resnet.py:236-270   -> oracle/uoc_oracle.py segnet_forward
                                                                                   (device agnostic: weights on the GPU)
  lib/utils/mean_shift.py:128-189 select_smart_seeds   -> select_seeds    (100 passes, one D2H sync per seed, :176)
  lib/utils/mean_shift.py:79-109  seed_hill_climbing_ball -> hill_climb   (W [m,n] materialised per update)
  lib/utils/mean_shift.py:41-76   connected_components -> label_seeds     (<= m tiny mm + host round trips)
  lib/utils/mean_shift.py:206-227 assignment + swap    -> assign_and_relabel
  lib/fcn/test_dataset.py:44-59   clustering_features  -> clustering_features (float32 CPU label maps)

Device shim (documented and minimal): the reference indexes CPU label tensors with CUDA masks /
indices at mean_shift.py:66,74 and :215, which current PyTorch rejects ("indices should be either on cpu or on the
same device as the indexed tensor"); here those masks / indices are moved with .cpu() first -- the cheapest change
that makes the unmodified algorithm run.  Everything else follows the reference's device placement: `cluster_labels`
and `selected_indices` live on the CPU, X / seeds / distances on the GPU.

Only bench.py's `torch_gpu_baseline` leg and tests/ import this file; results are checked against the CPU oracle in
tests/test_gpu_baselines.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

import uoc_oracle as O


def label_mode(values):
    return O.label_mode(values)


def select_seeds(X, num_seeds, first_index):
    """lib/utils/mean_shift.py:128-189 (cosine, fresh start) with X on the GPU."""
    n, d = X.shape
    selected = -1 * torch.ones(num_seeds, dtype=torch.long)                 # CPU (:140)
    seeds = torch.empty((num_seeds, d), device=X.device)                    # :144
    distances = torch.empty((n, num_seeds), device=X.device)                # :151
    selected[0] = int(first_index)
    seed = X[int(first_index), :]
    seeds[0, :] = seed
    distances[:, 0] = 0.5 * (1 - torch.mm(X, seed.unsqueeze(1))[:, 0])      # :162
    for i in range(1, num_seeds):
        nearest = torch.min(distances[:, :i], dim=1)[0]                     # :174 (re-reads i columns)
        idx = torch.argmax(nearest)                                         # :175
        selected[i] = idx                                                   # :176  D2H synchronisation
        seed = torch.index_select(X, 0, idx)[0, :]                          # :177
        seeds[i, :] = seed
        distances[:, i] = 0.5 * (1 - torch.mm(X, seed.unsqueeze(1))[:, 0])  # :184
    return seeds, selected


def hill_climb(X, Z, kappa, max_iters):
    """lib/utils/mean_shift.py:79-109 (cosine)."""
    for _ in range(max_iters):
        new_Z = Z.clone()                                                   # :93 (dead code in the reference, kept)
        W = torch.exp(kappa * torch.mm(Z, X.t()))                           # :26
        new_Z = torch.mm(W, X)                                              # :98
        Z = F.normalize(new_Z, p=2, dim=1)                                  # :107
    return Z


def label_seeds(Z, epsilon):
    """lib/utils/mean_shift.py:41-76 (cosine); labels on the CPU like the reference, masks moved with .cpu() (shim)."""
    n = Z.shape[0]
    K = 0
    labels = torch.ones(n, dtype=torch.long) * -1
    for i in range(n):
        if labels[i] == -1:
            distances = 0.5 * (1 - torch.mm(Z, Z[i:i + 1].t()))
            comp = (distances[:, 0] <= epsilon).cpu()                       # shim for :66 / :74
            if torch.unique(labels[comp]).shape[0] > 1:
                temp = labels[comp].numpy()
                temp = temp[temp != -1]
                label = torch.tensor(label_mode(temp))
            else:
                label = torch.tensor(K)
                K += 1
            labels[comp] = label
    return labels


def assign_and_relabel(X, Z, seed_labels):
    """lib/utils/mean_shift.py:206-227 (cosine)."""
    distances = 0.5 * (1 - torch.mm(X, Z.t()))                              # :211  [n, m] materialised
    closest = torch.argmin(distances, dim=1)                                # :214
    labels = seed_labels[closest.cpu()]                                     # :215 (shim) -> CPU int64 [n]
    num = len(torch.unique(seed_labels))
    count = torch.zeros(num, dtype=torch.long)
    for i in range(num):
        count[i] = (labels == i).sum()
    label_max = torch.argmax(count)
    if label_max != 0:
        index1 = labels == 0
        index2 = labels == label_max
        labels[index1] = label_max
        labels[index2] = 0
    return labels


def mean_shift_smart_init(X, kappa=O.KAPPA, num_seeds=O.NUM_SEEDS, max_iters=O.MAX_ITERS, first_index=None, return_all=False):
    n = X.shape[0]
    if first_index is None:
        first_index = np.random.randint(0, n)
    seeds, selected = select_seeds(X, num_seeds, first_index)
    Z = hill_climb(X, seeds, kappa, max_iters)
    seed_labels = label_seeds(Z, 2 * O.EMBEDDING_ALPHA)
    labels = assign_and_relabel(X, Z, seed_labels)
    if return_all:
        return labels, selected, seeds, Z, seed_labels
    return labels, selected


def clustering_features(features, num_seeds=O.NUM_SEEDS, first_indices=None):
    """lib/fcn/test_dataset.py:44-59: features on the GPU -> float32 CPU label maps."""
    N, C, H, W = features.shape
    out = torch.zeros((N, H, W))
    picked = []
    for j in range(N):
        X = features[j].view(C, -1).t()
        fi = None if first_indices is None else int(first_indices[j])
        labels, sel = mean_shift_smart_init(X, O.KAPPA, num_seeds, O.MAX_ITERS, fi)
        out[j] = labels.view(H, W)
        picked.append(sel)
    return out, picked


class TorchGpuSegNet(object):
    """The reference network on a CUDA device: same F.conv2d / F.batch_norm / F.interpolate / F.normalize calls as the
    oracle (uoc_oracle.segnet_forward), weights resident on the device, eval mode.  grad=True keeps autograd on like
    the reference does at inference (no torch.no_grad anywhere)."""

    def __init__(self, state_dict, device, grad=False):
        self.sd = {k: v.detach().float().to(device) for k, v in O.normalize_state_dict(state_dict).items()}
        self.grad = bool(grad)
        if self.grad:
            for k, v in self.sd.items():
                if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                    v.requires_grad_(True)

    def __call__(self, img, label=None, depth=None):
        if self.grad:
            return O.segnet_rgbd_add_forward(self.sd, img, depth)
        with torch.no_grad():
            return O.segnet_rgbd_add_forward(self.sd, img, depth)


def frame(net, img_host, xyz_host, first_index, device):
    """One frame the way lib/fcn/test_dataset.py:232-252 runs it (stage 1): H2D of the sample, network, .detach(),
    clustering_features (float32 CPU labels out)."""
    image = img_host.to(device)
    depth = xyz_host.to(device)
    features = net(image, None, depth).detach()
    out_label, _ = clustering_features(features, O.NUM_SEEDS, [first_index])
    return out_label
