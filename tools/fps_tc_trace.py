"""Debug: per-pass timeline of the screened seed selection kernel on a bench frame (UOC_FPS_TC_TRACE=<cta>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import mean_shift as MS, networks, synthetic

dev = torch.device("cuda:0")
H, W, D = 480, 640, 64
net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
img, depth = synthetic.rgbd_frame(H, W, seed=0)
feats = net(img.to(dev), None, depth.to(dev))
torch.cuda.synchronize()
for cta in [int(a) for a in sys.argv[1:]] or [0, 70]:
    os.environ["UOC_FPS_TC_TRACE"] = str(cta)
    MS.cluster_fields(feats, 100, 20.0, 10, [H * W // 5], epsilon=0.04)
    torch.cuda.synchronize()
