"""Debug: how many exact-distance rounds the screened seed selection runs per pass (UOC_FPS_TC_STATS), on the bench frame
(random-init backbone) and on a clustered synthetic field."""
import os, sys
os.environ["UOC_FPS_TC_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS, networks, synthetic

dev = torch.device("cuda:0")
H, W, D = 480, 640, 64
net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
img, depth = synthetic.rgbd_frame(H, W, seed=0)
feats = net(img.to(dev), None, depth.to(dev))
torch.cuda.synchronize()
sys.stderr.write("--- bench frame (random-init backbone)\n")
MS.cluster_fields(feats, 100, 20.0, 10, [H * W // 5], epsilon=0.04)
torch.cuda.synchronize()
f2, _ = synthetic.clustered_features(H, W, D, 6, 0.05, seed=0)
f2 = f2.to(dev)
MS.register_bf16_copy(f2, MS.pack_bf16(f2))
sys.stderr.write("--- clustered synthetic field (6 objects)\n")
MS.cluster_fields(f2, 100, 20.0, 10, [H * W // 5], epsilon=0.04)
torch.cuda.synchronize()
