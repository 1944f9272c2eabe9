"""The torch-GPU baseline (oracle/uoc_torch_gpu.py: the reference's PyTorch path restated for a CUDA device, with the
documented device shim) must compute what the CPU oracle computes -- otherwise bench.py's `torch_gpu_baseline` would time
something else than the reference's algorithm."""
import numpy as np
import pytest
import torch

import uoc_oracle as O
import uoc_torch_gpu as G
from unseenobjectclustering_b200 import networks as NW

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_torch_gpu_clustering_equals_cpu_oracle():
    feats, gt = O.synthetic_clustered_features(60, 80, 64, 4, 0.05, seed=3)
    want, sel_w = O.clustering_features(feats, 100, [17])
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        got, sel_g = G.clustering_features(feats.to(DEV), 100, [17])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved
    assert got.dtype == torch.float32 and got.device.type == "cpu"
    assert O.labels_equal_up_to_permutation(got.numpy(), want.numpy())
    assert O.labels_equal_up_to_permutation(got.numpy(), gt.numpy())
    assert int((sel_g[0] == sel_w[0]).sum()) >= 90          # cuBLAS vs MKL summation order: near-ties may differ


def test_torch_gpu_network_equals_cpu_oracle():
    sd = O.randomise_bn_(NW.random_state_dict(64, seed=2), 1002)
    img, xyz = O.synthetic_rgbd_frame(64, 96, seed=4)
    want = O.OracleSegNet(sd)(img, None, xyz)
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for grad in (False, True):
            got = G.TorchGpuSegNet(sd, DEV, grad=grad)(img.to(DEV), None, xyz.to(DEV)).detach().cpu()
            assert float((1.0 - (got * want).sum(1)).abs().max()) < 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = saved
