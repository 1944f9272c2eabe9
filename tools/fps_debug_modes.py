"""Measurement: sampling kernel time with / without the inter-CTA exchange for several shared-memory residencies."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectclustering_b200 import mean_shift as MS, synthetic
feats, _ = synthetic.clustered_features(480, 640, 64, 6, 0.05, 0)
X = feats.cuda()[0].view(64, -1).t()
for smem in (0, 48, 96, 150, 200):
    for mode in (0, 1):
        os.environ["UOC_FPS_DEBUG_MODE"] = str(mode)
        os.environ["UOC_FPS_SMEM_KB"] = str(smem)
        for rep in range(3):
            MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(10):
            MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
        e1.record(); torch.cuda.synchronize()
        print("resident %3d KB/SM, %s: %.3f ms per sampling call" % (smem, "no exchange" if mode else "normal     ", e0.elapsed_time(e1) / 10))
