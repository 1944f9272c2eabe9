"""GPU parity of the input preparation through the C ABI: bit-exact against the golden outputs of the
reference's own read_sample and against the oracle restatement on full-size random frames."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu

import uoc_oracle as O                                              # noqa: E402
from unseenobjectclustering_b200 import input_prep as IP            # noqa: E402
from unseenobjectclustering_b200 import _lib                        # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "input_prep.npz")
DEV = torch.device("cuda:0")


def _cam(g):
    return {"fx": float(g["fx"]), "fy": float(g["fy"]), "x_offset": float(g["x_offset"]), "y_offset": float(g["y_offset"])}


def test_matches_reference_golden_bit_exact():
    g = np.load(GOLDEN)
    for k in range(int(g["cases"])):
        image, xyz = IP.prepare_inputs(g["im%d" % k], g["depth%d" % k], _cam(g), device=DEV)
        assert image.dtype == torch.float32 and xyz.dtype == torch.float32 and image.is_cuda
        assert np.array_equal(image.cpu().numpy(), g["image_color%d" % k])
        assert np.array_equal(xyz.cpu().numpy(), g["xyz%d" % k])


@pytest.mark.parametrize("H,W,N", [(480, 640, 1), (37, 53, 3), (720, 960, 1)])
def test_matches_oracle_bit_exact_full_range(H, W, N):
    rng = np.random.RandomState(H + N)
    im = rng.randint(0, 256, (N, H, W, 3)).astype(np.uint8)
    dp = rng.randint(0, 65536, (N, H, W)).astype(np.uint16)         # the full uint16 range, incl. values >= 32768 and zeros
    dp[:, ::7, ::5] = 0
    cam = {"fx": 612.937, "fy": 613.173, "x_offset": 322.549, "y_offset": 248.158}
    image, xyz = IP.prepare_inputs(im, dp, cam, device=DEV)
    for n in range(N):
        io, xo = O.read_sample_arrays(im[n], dp[n], cam)
        assert np.array_equal(image[n].cpu().numpy(), io[0].numpy())
        assert np.array_equal(xyz[n].cpu().numpy(), xo[0].numpy())


def test_single_modalities_and_compute_xyz_mirror():
    rng = np.random.RandomState(3)
    im = rng.randint(0, 256, (24, 40, 3)).astype(np.uint8)
    dp = rng.randint(0, 4000, (24, 40)).astype(np.uint16)
    cam = {"fx": 500.0, "fy": 505.5, "x_offset": 19.25, "y_offset": 11.75}
    io, xo = O.read_sample_arrays(im, dp, cam)
    image, none = IP.prepare_inputs(im, None, cam, device=DEV)
    assert none is None and np.array_equal(image.cpu().numpy(), io.numpy())
    none, xyz = IP.prepare_inputs(None, dp, cam, device=DEV)
    assert none is None and np.array_equal(xyz.cpu().numpy(), xo.numpy())
    depth_m = torch.from_numpy(dp.astype(np.float32) / 1000.0).to(DEV)
    got = IP.compute_xyz(depth_m, cam["fx"], cam["fy"], cam["x_offset"], cam["y_offset"], 24, 40)
    want = O.compute_xyz(dp.astype(np.float32) / 1000.0, cam["fx"], cam["fy"], cam["x_offset"], cam["y_offset"], 24, 40)
    assert tuple(got.shape) == (24, 40, 3) and np.array_equal(got.cpu().numpy(), want)


def test_errors_are_loud():
    with pytest.raises(_lib.UocError):
        IP.prepare_inputs(None, None, {})
    with pytest.raises(_lib.UocError):
        IP.prepare_inputs(np.zeros((4, 4, 3), np.uint8), np.zeros((5, 4), np.uint16), {"fx": 1.0, "fy": 1.0}, device=DEV)
    with pytest.raises(_lib.UocError):
        IP.prepare_inputs(np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4), np.uint16), {"fx": 0.0, "fy": 1.0}, device=DEV)
    with pytest.raises(_lib.UocError):
        IP.compute_xyz(torch.zeros(4, 4), 1.0, 1.0, 0.0, 0.0, 4, 4)
