"""The hot path once per batch size, for ncu launch lists / captures: warm-up, then ONE batch-1 frame and ONE batch-4 step
(eager launches: every kernel appears under its own name).  usage: python tools/profile_frame.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS, networks, synthetic
dev = torch.device("cuda:0")
net = networks.seg_resnet34_8s_embedding(2, 64, networks.random_state_dict(64, seed=0)).to(dev)
for B in ([int(v) for v in sys.argv[1:]] or [1, 4]):
    img, xyz = synthetic.rgbd_frame(480, 640, seed=0, batch=B)
    img, xyz = img.to(dev), xyz.to(dev)
    for rep in range(3):                       # 2 warm-up passes + the profiled one
        f, xb = net.forward_ex(img, None, xyz)
        lab, sel = MS.cluster_fields(f, 100, first_indices=[1000 + 17 * i for i in range(B)], x_bf16=xb)
        torch.cuda.synchronize()
    print("B", B, "labels", int(lab.max()), flush=True)
