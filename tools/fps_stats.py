"""Sampling kernel statistics (UOC_FPS_STATS=1): how many passes end with an exchange, clocks per pass, on the bench frame
(embedding of the random-init backbone) and on the clustered config-2 field."""
import os, sys
os.environ["UOC_FPS_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS, networks, synthetic
dev = torch.device("cuda:0")
net = networks.seg_resnet34_8s_embedding(2, 64, networks.random_state_dict(64, seed=0)).to(dev)
img, xyz = synthetic.rgbd_frame(480, 640, seed=0)
f, xb = net.forward_ex(img.to(dev), None, xyz.to(dev))
for rep in range(2):
    print("bench frame", flush=True)
    MS.cluster_fields(f, 100, first_indices=[1000 + rep], x_bf16=xb)
    torch.cuda.synchronize()
fc, _ = synthetic.clustered_features(480, 640, 64, 6, 0.05, seed=0)
fc = fc.to(dev)
xc = MS.pack_bf16(fc)
for rep in range(2):
    print("clustered field", flush=True)
    MS.cluster_fields(fc, 100, first_indices=[2000 + rep], x_bf16=xc)
    torch.cuda.synchronize()
