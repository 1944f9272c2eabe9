"""shim.main() on a tool with the reference's layout, CPU half: tests/ref_layout is a stand-in tree (module names and call
sequence of the reference's tools/test_images.py, every function of the path RAISES unless rebound).  Here the tool is run
under the shim up to the point where the networks are built; the GPU half (tests/test_gpu_zz_shim_tool.py) lets it segment
frames.  Also: the same run against the real reference tree where it exists (this container)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

TOOL = os.path.join(ROOT, "tests", "ref_layout", "tools", "segment_images.py")


def _run(args, **kw):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    return subprocess.run([sys.executable, "-m", "unseenobjectclustering_b200.shim"] + args, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600, **kw)


def _frames(tmp_path, n=1, H=48, W=64):
    import cv2
    import json
    rng = np.random.RandomState(0)
    for i in range(n):
        cv2.imwrite(str(tmp_path / ("%03d-color.png" % i)), rng.randint(0, 256, (H, W, 3)).astype(np.uint8))
        cv2.imwrite(str(tmp_path / ("%03d-depth.png" % i)), rng.randint(400, 1500, (H, W)).astype(np.uint16))
    with open(tmp_path / "camera_params.json", "w") as f:
        json.dump({"fx": 60.0, "fy": 61.0, "x_offset": 31.5, "y_offset": 23.5}, f)


def test_stand_in_tool_fails_without_the_shim(tmp_path):
    """The stand-in's own functions raise: a green run under the shim proves the rebinding, not the stand-in."""
    from unseenobjectclustering_b200 import networks as N
    _frames(tmp_path)
    torch.save(N.random_state_dict(64, seed=0), str(tmp_path / "ckpt.pth"))
    r = subprocess.run([sys.executable, TOOL, "--pretrained", str(tmp_path / "ckpt.pth"), "--imgdir", str(tmp_path),
                        "--stop-after-build"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "did not rebind" in r.stderr


@pytest.mark.parametrize("variant", ["rgbd_add", "color"])
def test_tool_under_shim_builds_this_packages_networks(tmp_path, variant):
    from unseenobjectclustering_b200 import networks as N
    _frames(tmp_path)
    torch.save({"module." + k: v for k, v in N.random_state_dict(64, seed=0).items()}, str(tmp_path / "ckpt.pth"))
    args = [TOOL, "--pretrained", str(tmp_path / "ckpt.pth"), "--pretrained_crop", str(tmp_path / "ckpt.pth"),
            "--imgdir", str(tmp_path), "--stop-after-build"]
    if variant == "color":                                   # cfg_from_file() runs after the imports: read live
        (tmp_path / "color.yml").write_text("INPUT: COLOR\nTRAIN:\n  FUSION_TYPE: add\n")
        args += ["--cfg", str(tmp_path / "color.yml")]
    r = _run(args)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "shim patched:" in r.stdout and "fcn.test_dataset.test_sample" in r.stdout and "cfg (" in r.stdout
    want = "built: SEGNET_B200 %s 64 test_sample from unseenobjectclustering_b200.test_dataset" % ("COLOR" if variant == "color" else "RGBD")
    assert want in r.stdout, r.stdout[-2000:]


def test_unmodified_reference_tool_under_shim_reaches_this_packages_forward(tmp_path):
    """The reference's own tools/test_images.py, unmodified, under shim.main() on its own demo frames (this container only):
    it parses its arguments, reads its yaml cfg, loads the checkpoint, builds the network through the rebound factory, wraps it
    in DataParallel, reads a demo frame with its own read_sample and calls the rebound test_sample, which calls
    SEGNET_B200.forward -- and on this CPU-only box that fails LOUDLY (there is no CPU path).  On a B200 the same call
    sequence is exercised by tests/test_gpu_zz_shim_tool.py with a stand-in tree (the reference tree cannot travel)."""
    import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present (GPU box)")
    if torch.cuda.is_available():
        pytest.skip("CPU half; the GPU half is tests/test_gpu_zz_shim_tool.py")
    from unseenobjectclustering_b200 import networks as N
    torch.save(N.random_state_dict(64, seed=0), str(tmp_path / "ckpt.pth"))
    ref = rh.REF_ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_ref_tool_under_shim.py"), "--cpu-cuda-identity",
                        os.path.join(ref, "tools", "test_images.py"), "--gpu", "0", "--network", "seg_resnet34_8s_embedding",
                        "--pretrained", str(tmp_path / "ckpt.pth"), "--imgdir", os.path.join(ref, "data", "demo"),
                        "--cfg", os.path.join(ref, "experiments", "cfgs", "seg_resnet34_8s_embedding_cosine_rgbd_add_tabletop.yml")],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode != 0
    assert "shim patched:" in r.stdout and "=> using pre-trained network" in r.stdout and "-color.png" in r.stdout
    tb = r.stderr
    assert "tools/test_images.py" in tb and "out_label, out_label_refined = test_sample(sample, network, network_crop)" in tb
    assert "unseenobjectclustering_b200/test_dataset.py" in tb and "unseenobjectclustering_b200/networks.py" in tb
    assert "there is no CPU path in this package" in tb.strip().splitlines()[-1]


def test_stand_in_tool_under_shim_reaches_this_packages_forward_on_cpu(tmp_path):
    """The whole call sequence of the stand-in tool (factory -> .cuda -> DataParallel -> .eval -> cv2 frames -> test_sample)
    with Tensor.cuda() as a no-op: it must arrive in SEGNET_B200.forward and fail loudly there (no CPU path)."""
    if torch.cuda.is_available():
        pytest.skip("CPU half; the GPU half is tests/test_gpu_zz_shim_tool.py")
    from unseenobjectclustering_b200 import networks as N
    _frames(tmp_path)
    torch.save(N.random_state_dict(64, seed=0), str(tmp_path / "ckpt.pth"))
    code = ("import sys; sys.path[:0] = [%r, %r]; import ref_harness as rh; from unseenobjectclustering_b200 import shim\n"
            "with rh.cpu_cuda_identity():\n    shim.main(sys.argv[1:])\n" % (ROOT, os.path.join(ROOT, "oracle")))
    r = subprocess.run([sys.executable, "-c", code, TOOL, "--pretrained", str(tmp_path / "ckpt.pth"), "--pretrained_crop",
                        str(tmp_path / "ckpt.pth"), "--imgdir", str(tmp_path)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "000-color.png" in r.stdout
    assert "segment_images.py" in r.stderr and "unseenobjectclustering_b200/test_dataset.py" in r.stderr
    assert "there is no CPU path in this package" in r.stderr.strip().splitlines()[-1]
