import os, sys, torch
sys.path.insert(0, '/root/repo')
from unseenobjectclustering_b200 import mean_shift as MS, synthetic
feats, _ = synthetic.clustered_features(480, 640, 64, 6, 0.05, 0)
X = feats.cuda()[0].view(64, -1).t()
for mode in (0, 1, 2):
    os.environ["UOC_FPS_DEBUG_MODE"] = str(mode)
    for rep in range(3):
        MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(5):
        MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
    e1.record(); torch.cuda.synchronize()
    print("debug mode %d (0 normal, 1 no exchange, 2 no streaming loop): %.3f ms per sampling call" % (mode, e0.elapsed_time(e1) / 5))
