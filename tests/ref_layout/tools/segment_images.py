#!/usr/bin/env python3
"""Stand-in tool with the call sequence of the reference's tools/test_images.py:136-223 (see ../README.md).  It is meant to be
run as  python -m unseenobjectclustering_b200.shim tests/ref_layout/tools/segment_images.py --pretrained ... --imgdir ...
and knows nothing about this package: it imports `networks`, `fcn.config`, `fcn.test_dataset` from its own lib/."""
import argparse
import glob
import json
import os
import sys

import cv2
import numpy as np
import torch
import torch.backends.cudnn as cudnn

import _init_paths  # noqa: F401
from fcn.test_dataset import test_sample            # bound at import time, like the reference tool does
from fcn.config import cfg, cfg_from_file
import networks


def options():
    p = argparse.ArgumentParser(description="stand-in for the reference's tools/test_images.py (same option names)")
    p.add_argument('--gpu', dest='gpu_id', default=0, type=int)
    p.add_argument('--pretrained', default=None, type=str)
    p.add_argument('--pretrained_crop', default=None, type=str)
    p.add_argument('--cfg', dest='cfg_file', default=None, type=str)
    p.add_argument('--imgdir', default=None, type=str)
    p.add_argument('--color', dest='color_name', default='*color.png', type=str)
    p.add_argument('--depth', dest='depth_name', default='*depth.png', type=str)
    p.add_argument('--network', dest='network_name', default='seg_resnet34_8s_embedding', type=str)
    p.add_argument('--out', default=None, type=str, help='npz that receives the label maps')
    p.add_argument('--stop-after-build', action='store_true', help='CPU check: build the networks, report, stop')
    return p.parse_args()


def back_project(depth, cam):
    """metric depth [H,W] -> XYZ [H,W,3] with the pinhole model of camera_params.json"""
    rows, cols = depth.shape
    u = np.tile(np.arange(cols, dtype=np.float32)[None, :], (rows, 1))
    v = np.tile(np.arange(rows, dtype=np.float32)[:, None], (1, cols))
    return np.stack([(u - cam['x_offset']) * depth / cam['fx'], (v - cam['y_offset']) * depth / cam['fy'], depth],
                    axis=-1).astype(np.float32)


def load_frame(color_png, depth_png, cam):
    """What the reference tool's read_sample hands to test_sample: BGR / 255 - PIXEL_MEANS / 255 as [1,3,H,W]; depth in
    millimetres -> metres -> XYZ as [1,3,H,W] (only for the DEPTH / RGBD inputs)."""
    bgr = torch.from_numpy(cv2.imread(color_png)) / 255.0
    bgr -= torch.tensor(cfg.PIXEL_MEANS / 255.0).float()
    frame = {'image_color': bgr.permute(2, 0, 1).unsqueeze(0)}
    if cfg.INPUT in ('DEPTH', 'RGBD'):
        metres = cv2.imread(depth_png, cv2.IMREAD_ANYDEPTH).astype(np.float32) / 1000.0
        frame['depth'] = torch.from_numpy(back_project(metres, cam)).permute(2, 0, 1).unsqueeze(0)
    return frame


def construct(name, checkpoint):
    """factory call of the reference tool: networks.__dict__[name](num_classes, num_units, checkpoint dict)"""
    return networks.__dict__[name](2, cfg.TRAIN.NUM_UNITS, torch.load(checkpoint))


def deploy(net):
    """.cuda -> DataParallel on the one device -> .cuda -> eval, the wrapping the reference tool applies"""
    net = net.cuda(device=cfg.device)
    net = torch.nn.DataParallel(net, device_ids=[cfg.gpu_id]).cuda(device=cfg.device)
    net.eval()
    return net


def main():
    opts = options()
    if opts.cfg_file:
        cfg_from_file(opts.cfg_file)                     # AFTER the imports: the factories must read cfg at construction
    np.random.seed(cfg.RNG_SEED)
    cfg.gpu_id, cfg.MODE = 0, 'TEST'
    cfg.device = torch.device('cuda:%d' % cfg.gpu_id)
    colors = sorted(glob.glob(os.path.join(opts.imgdir, opts.color_name)))
    depths = sorted(glob.glob(os.path.join(opts.imgdir, opts.depth_name)))
    with open(os.path.join(opts.imgdir, 'camera_params.json')) as fh:
        cam = json.load(fh)
    if not opts.pretrained:
        sys.exit("no pretrained network specified")
    net = construct(opts.network_name, opts.pretrained)
    net_crop = construct(opts.network_name, opts.pretrained_crop) if opts.pretrained_crop else None
    if opts.stop_after_build:
        print("built:", type(net).__name__, getattr(net, "input_type", "?"), getattr(net, "feature_dim", "?"),
              "test_sample from", test_sample.__module__)
        return
    cudnn.benchmark = True
    net = deploy(net)
    if net_crop is not None:
        net_crop = deploy(net_crop)
    maps = {}
    for k, (c, d) in enumerate(zip(colors, depths)):
        print(c)
        first_stage, second_stage = test_sample(load_frame(c, d, cam), net, net_crop)
        maps["out_label_%d" % k] = first_stage.numpy()
        if second_stage is not None:
            maps["out_label_refined_%d" % k] = second_stage.numpy()
    if opts.out:
        np.savez(opts.out, **maps)
    print("segmented %d frames" % len(colors))


if __name__ == '__main__':
    main()
