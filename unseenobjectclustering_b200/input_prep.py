"""Input preparation on the device: host mirror of the reference's `read_sample` / `compute_xyz`
(tools/test_images.py:96-135; the same arithmetic in ros/test_images_segmentation.py:38-44,146-159).

The reference builds both network inputs on the CPU in fp32 (7.4 MB per 640x480 frame to upload); here the raw frame
(uint8 BGR + uint16 depth, 1.5 MB) is uploaded and one kernel writes both `[1,3,H,W]` tensors, bit-identical to the
reference's numpy / torch arithmetic (uoc_prepare_inputs, csrc/input_prep.cu).  No CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib

# fcn/config.py:376
PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])


def _means_over_255(pixel_means=None):
    pm = PIXEL_MEANS if pixel_means is None else np.asarray(pixel_means, dtype=np.float64)
    # read_sample: torch.tensor(cfg.PIXEL_MEANS / 255.0).float()  -- float64 division, then rounded to float32
    v = (pm.reshape(-1) / 255.0).astype(np.float32)
    if v.shape[0] != 3:
        raise _lib.UocError("PIXEL_MEANS must have 3 entries")
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def _device_tensor(a, dtype, device):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if t.dtype != dtype:
        raise _lib.UocError("expected %s, got %s" % (dtype, t.dtype))
    return t.to(device, non_blocking=True).contiguous()


def prepare_inputs(im_bgr, depth_raw, camera_params, device=None, pixel_means=None, depth_divisor=1000.0, out=None):
    """im_bgr: [H,W,3] or [N,H,W,3] uint8 (cv2.imread order) or None; depth_raw: [H,W] or [N,H,W] uint16 (cv2.IMREAD_ANYDEPTH;
    int16 storage of the same bits is accepted) or None; camera_params: dict with fx, fy, x_offset, y_offset.
    Returns (image_color [N,3,H,W] fp32 CUDA or None, depth [N,3,H,W] fp32 CUDA or None) -- the values of
    sample['image_color'] / sample['depth'] of tools/test_images.py:105-135.  out: optional (image, xyz) tensors to fill."""
    if im_bgr is None and depth_raw is None:
        raise _lib.UocError("neither colour nor depth given")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.UocError("input preparation runs on a CUDA device: there is no CPU path in this package")
    lib = _lib.load()
    im_t = dp_t = None
    shape = None
    if im_bgr is not None:
        im_t = _device_tensor(im_bgr, torch.uint8, dev)
        if im_t.dim() == 3:
            im_t = im_t.unsqueeze(0)
        if im_t.dim() != 4 or im_t.shape[3] != 3:
            raise _lib.UocError("colour image must be [H,W,3] or [N,H,W,3] uint8")
        shape = tuple(im_t.shape[:3])
    if depth_raw is not None:
        if isinstance(depth_raw, np.ndarray) and depth_raw.dtype == np.uint16:
            depth_raw = depth_raw.view(np.int16)          # torch has no uint16 arithmetic; the kernel reads the bits as uint16
        dp_t = _device_tensor(depth_raw, torch.int16, dev)
        if dp_t.dim() == 2:
            dp_t = dp_t.unsqueeze(0)
        if dp_t.dim() != 3 or (shape is not None and tuple(dp_t.shape) != shape):
            raise _lib.UocError("depth must be [H,W] or [N,H,W] uint16 with the colour image's size")
        shape = tuple(dp_t.shape)
    N, H, W = shape
    image = xyz = None
    if out is not None:
        image, xyz = out
    if im_t is not None and image is None:
        image = torch.empty((N, 3, H, W), dtype=torch.float32, device=dev)
    if dp_t is not None and xyz is None:
        xyz = torch.empty((N, 3, H, W), dtype=torch.float32, device=dev)
    for t in (image if im_t is not None else None, xyz if dp_t is not None else None):
        if t is not None and (tuple(t.shape) != (N, 3, H, W) or t.dtype != torch.float32 or not t.is_contiguous()):
            raise _lib.UocError("output tensors must be contiguous float32 [N,3,H,W]")
    cp = camera_params or {}
    fx, fy = float(cp.get("fx", 1.0)), float(cp.get("fy", 1.0))
    px, py = float(cp.get("x_offset", 0.0)), float(cp.get("y_offset", 0.0))
    with torch.cuda.device(dev):
        st = lib.uoc_prepare_inputs(_lib.ptr(im_t), _lib.ptr(dp_t), N, H, W, fx, fy, px, py,
                                    ctypes.cast(_means_over_255(pixel_means), ctypes.c_void_p), float(depth_divisor),
                                    _lib.ptr(image) if im_t is not None else None,
                                    _lib.ptr(xyz) if dp_t is not None else None, _lib.stream_ptr(dev))
        _lib.check(st, "uoc_prepare_inputs")
    return (image if im_t is not None else None), (xyz if dp_t is not None else None)


def compute_xyz(depth_img, fx, fy, px, py, height, width):
    """tools/test_images.py:96-102 on the device: metric depth [H,W] fp32 (CUDA tensor) -> [H,W,3] fp32 view (x, y, z)."""
    if not (torch.is_tensor(depth_img) and depth_img.is_cuda):
        raise _lib.UocError("depth must be a CUDA tensor: there is no CPU path in this package")
    if depth_img.dtype != torch.float32 or tuple(depth_img.shape) != (height, width):
        raise _lib.UocError("depth must be float32 [height, width]")
    d = depth_img.contiguous()
    out = torch.empty((1, 3, height, width), dtype=torch.float32, device=d.device)
    lib = _lib.load()
    with torch.cuda.device(d.device):
        st = lib.uoc_compute_xyz(_lib.ptr(d), 1, height, width, float(fx), float(fy), float(px), float(py), _lib.ptr(out),
                                 _lib.stream_ptr(d.device))
        _lib.check(st, "uoc_compute_xyz")
    return out[0].permute(1, 2, 0)


def read_sample(filename_color, filename_depth, camera_params, input_mode="RGBD", device=None):
    """tools/test_images.py:105-135: same files, same dict -- but the tensors are built on (and stay on) the device."""
    import cv2
    im = cv2.imread(filename_color)
    depth_img = None
    if input_mode in ("DEPTH", "RGBD"):
        depth_img = cv2.imread(filename_depth, cv2.IMREAD_ANYDEPTH)
    image, xyz = prepare_inputs(im, depth_img, camera_params, device=device)
    sample = {"image_color": image}
    if xyz is not None:
        sample["depth"] = xyz
    return sample
