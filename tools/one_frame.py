"""One 640x480 frame through the public path (backbone + clustering), twice: the target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS, networks, synthetic

dev = torch.device("cuda:0")
H, W, D = 480, 640, 64
net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
img, depth = synthetic.rgbd_frame(H, W, seed=0)
img, depth = img.to(dev), depth.to(dev)
for rep in range(2):
    feats = net(img, None, depth)
    labels, sel = MS.cluster_fields(feats, 100, 20.0, 10, [H * W // 5], epsilon=0.04)
    torch.cuda.synchronize()
print("labels", int(labels.max()) + 1, "clusters")
