"""Where the time of one two-stage frame goes (wall clock with a device synchronise after every stage)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import networks, synthetic, test_dataset as TD

dev = torch.device("cuda:0")
H, W, D = 480, 640, 64
net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
net_crop = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=1)).to(dev)
feats1, _ = synthetic.clustered_features(H, W, D, 6, 0.05, seed=0)
feats1 = feats1.to(dev)
g = torch.Generator().manual_seed(5)
centres = torch.nn.functional.normalize(torch.randn(2, D, generator=g), dim=1).to(dev)
noise = (0.05 * torch.randn(8, D, 224, 224, generator=g)).to(dev)
img, xyz = synthetic.rgbd_frame(H, W, seed=0)
img, xyz = img.pin_memory(), xyz.pin_memory()
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3
    return t1


reps = 8
for rep in range(reps + 2):
    if rep == 2:
        acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    image, depth = img.cuda(), xyz.cuda()
    t = tick("h2d", t)
    net(image, None, depth)
    t = tick("backbone", t)
    labels, _ = TD.clustering_features_device(feats1, 100, [1000 + rep])
    t = tick("clustering", t)
    labels = labels.view(1, H, W)
    labels = TD._filter_labels_depth_device(labels, depth, 0.8)
    out_label = labels.to(torch.float32).cpu()
    t = tick("depth filter + d2h", t)
    rgb_crop, out_label_crop, rois, depth_crop = TD.crop_rois(image, labels, depth)
    t = tick("crop_rois", t)
    net_crop(rgb_crop, out_label_crop, depth_crop)
    t = tick("crop backbone (K=%d)" % rgb_crop.shape[0], t)
    ids = (out_label_crop > 0).long()
    f = torch.nn.functional.normalize(centres[ids].permute(0, 3, 1, 2) + noise[:rgb_crop.shape[0]], dim=1).contiguous()
    torch.cuda.synchronize()
    t = time.perf_counter()
    labels_crop, _ = TD.clustering_features_device(f, 100, [77 + k for k in range(8)])
    t = tick("crop clustering", t)
    labels_crop = labels_crop.view(-1, 224, 224).to(torch.float32)
    refined, _ = TD.match_label_crop(out_label, labels_crop, out_label_crop, rois, depth_crop)
    t = tick("match_label_crop", t)
tot = 0.0
for k, v in acc.items():
    print("%-28s %.3f ms" % (k, v / reps))
    tot += v / reps
print("%-28s %.3f ms" % ("sum", tot))
