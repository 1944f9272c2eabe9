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


def parse_args():
    p = argparse.ArgumentParser()
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


def compute_xyz(depth, fx, fy, px, py):
    h, w = depth.shape
    ix = np.tile(np.arange(w, dtype=np.float32)[None, :], (h, 1))
    iy = np.tile(np.arange(h, dtype=np.float32)[:, None], (1, w))
    return np.stack([(ix - px) * depth / fx, (iy - py) * depth / fy, depth], axis=-1).astype(np.float32)


def read_sample(file_color, file_depth, camera_params):
    """color: BGR / 255 - PIXEL_MEANS / 255, [1,3,H,W]; depth: millimetres -> metres -> XYZ [1,3,H,W]"""
    im = cv2.imread(file_color)
    depth = cv2.imread(file_depth, cv2.IMREAD_ANYDEPTH).astype(np.float32) / 1000.0
    xyz = compute_xyz(depth, camera_params['fx'], camera_params['fy'], camera_params['x_offset'], camera_params['y_offset'])
    im_tensor = torch.from_numpy(im) / 255.0
    im_tensor -= torch.tensor(cfg.PIXEL_MEANS / 255.0).float()
    sample = {'image_color': im_tensor.permute(2, 0, 1).unsqueeze(0)}
    if cfg.INPUT in ('DEPTH', 'RGBD'):
        sample['depth'] = torch.from_numpy(xyz).permute(2, 0, 1).unsqueeze(0)
    return sample


if __name__ == '__main__':
    args = parse_args()
    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)                     # AFTER the imports: the factories must read cfg at construction
    np.random.seed(cfg.RNG_SEED)
    cfg.gpu_id = 0
    cfg.device = torch.device('cuda:{:d}'.format(cfg.gpu_id))
    cfg.MODE = 'TEST'
    images_color = sorted(glob.glob(os.path.join(args.imgdir, args.color_name)))
    images_depth = sorted(glob.glob(os.path.join(args.imgdir, args.depth_name)))
    with open(os.path.join(args.imgdir, 'camera_params.json')) as f:
        camera_params = json.load(f)
    if not args.pretrained:
        sys.exit("no pretrained network specified")
    network_data = torch.load(args.pretrained)
    network = networks.__dict__[args.network_name](2, cfg.TRAIN.NUM_UNITS, network_data)
    network_crop = None
    if args.pretrained_crop:
        network_crop = networks.__dict__[args.network_name](2, cfg.TRAIN.NUM_UNITS, torch.load(args.pretrained_crop))
    if args.stop_after_build:
        print("built:", type(network).__name__, getattr(network, "input_type", "?"), getattr(network, "feature_dim", "?"),
              "test_sample from", test_sample.__module__)
        sys.exit(0)
    network = network.cuda(device=cfg.device)
    network = torch.nn.DataParallel(network, device_ids=[cfg.gpu_id]).cuda(device=cfg.device)
    cudnn.benchmark = True
    network.eval()
    if network_crop is not None:
        network_crop = network_crop.cuda(device=cfg.device)
        network_crop = torch.nn.DataParallel(network_crop, device_ids=[cfg.gpu_id]).cuda(device=cfg.device)
        network_crop.eval()
    results = {}
    for i in range(len(images_color)):
        print(images_color[i])
        sample = read_sample(images_color[i], images_depth[i], camera_params)
        out_label, out_label_refined = test_sample(sample, network, network_crop)
        results["out_label_%d" % i] = out_label.numpy()
        if out_label_refined is not None:
            results["out_label_refined_%d" % i] = out_label_refined.numpy()
    if args.out:
        np.savez(args.out, **results)
    print("segmented %d frames" % len(images_color))
