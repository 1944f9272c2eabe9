"""A tool with the reference's layout and call sequence, run under `python -m unseenobjectclustering_b200.shim` ON THE GPU
(VERDICT round 1, item 5).  The reference tree cannot travel to the GPU box, so the tool is the stand-in of
tests/ref_layout (module names, import order and call sequence of tools/test_images.py:136-223; every function of the path
in its lib/ raises unless shim.install() rebound it).  The label maps the tool writes must equal, bit for bit, what this
package's test_sample returns in-process for the same frames, checkpoints and NumPy seed.  (Named zz: runs last.)"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import uoc_oracle as O
from conftest import ROOT
from unseenobjectclustering_b200 import networks as NW
from unseenobjectclustering_b200 import test_dataset as TD

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOOL = os.path.join(ROOT, "tests", "ref_layout", "tools", "segment_images.py")
CAM = {"fx": 150.0, "fy": 151.0, "x_offset": 79.5, "y_offset": 59.5}


def _raw_frame(H, W, num_objects, seed):
    """uint8 BGR + uint16 depth (mm) of a table-top-like scene: rectangles of constant colour standing 10-35 cm in front of a
    background plane at 1.2 m; depth is missing on the right 70 % of the frame."""
    rng = np.random.RandomState(seed)
    im = np.empty((H, W, 3), np.uint8)
    im[:] = rng.randint(0, 256, 3)
    dp = np.full((H, W), 1200, np.uint16)
    for _ in range(num_objects):
        h, w = rng.randint(H // 6, H // 3), rng.randint(W // 6, W // 3)
        y, x = rng.randint(0, H - h), rng.randint(0, W - w)
        im[y:y + h, x:x + w] = rng.randint(0, 256, 3)
        dp[y:y + h, x:x + w] = 1200 - rng.randint(100, 350)
    dp[:, int(W * 0.3):] = 0       # no depth on the right 70 %: the depth filter (test_dataset.py:183-198) drops the clusters there,
    return im, dp                  # which keeps the number of crops of the second stage moderate (10-14 with these weights)


def test_reference_layout_tool_under_shim_on_the_gpu(tmp_path):
    import cv2
    H, W, F = 120, 160, 3
    raws = [_raw_frame(H, W, 4, 40 + i) for i in range(F)]
    for i, (im, dp) in enumerate(raws):
        assert cv2.imwrite(str(tmp_path / ("%06d-color.png" % i)), im)
        assert cv2.imwrite(str(tmp_path / ("%06d-depth.png" % i)), dp)
    with open(tmp_path / "camera_params.json", "w") as f:
        json.dump(CAM, f)
    img0, xyz0 = O.read_sample_arrays(raws[0][0], raws[0][1], CAM)
    sd = O.calibrated_state_dict_(NW.random_state_dict(64, seed=0), img0, xyz0)          # embeddings that do not collapse
    sd_crop = NW.random_state_dict(64, seed=1)
    torch.save({"module." + k: v for k, v in sd.items()}, str(tmp_path / "ckpt.pth"))      # DataParallel-style keys
    torch.save(sd_crop, str(tmp_path / "ckpt_crop.pth"))
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-m", "unseenobjectclustering_b200.shim", TOOL, "--gpu", "0",
                        "--network", "seg_resnet34_8s_embedding", "--pretrained", str(tmp_path / "ckpt.pth"),
                        "--pretrained_crop", str(tmp_path / "ckpt_crop.pth"), "--imgdir", str(tmp_path),
                        "--out", str(tmp_path / "out.npz")], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "shim patched:" in r.stdout and "segmented %d frames" % F in r.stdout
    got = np.load(str(tmp_path / "out.npz"))
    # the same frames, checkpoints and NumPy stream through this package in-process
    net = NW.seg_resnet34_8s_embedding(2, 64, sd).to(DEV)
    net_crop = NW.seg_resnet34_8s_embedding(2, 64, sd_crop).to(DEV)
    np.random.seed(3)                                                                      # cfg.RNG_SEED of the stand-in cfg
    clusters = []
    for i, (im, dp) in enumerate(raws):
        img, xyz = O.read_sample_arrays(im, dp, CAM)
        out_label, refined = TD.test_sample({'image_color': img, 'depth': xyz}, net, net_crop)
        assert got["out_label_%d" % i].dtype == np.float32 and got["out_label_%d" % i].shape == (1, H, W)
        assert np.array_equal(got["out_label_%d" % i], out_label.numpy()), i
        assert (("out_label_refined_%d" % i) in got.files) == (refined is not None)
        if refined is not None:
            assert np.array_equal(got["out_label_refined_%d" % i], refined.numpy()), i
        clusters.append(int(len(np.unique(out_label.numpy()))))
    print("clusters per frame under the shim:", clusters)
