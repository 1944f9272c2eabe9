"""Debug: per-pass phase timing of the sampling kernel (UOC_FPS_TRACE=<cta>) on a 640x480x64 field."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectclustering_b200 import mean_shift as MS, synthetic

feats, _ = synthetic.clustered_features(480, 640, 64, 6, 0.05, 0)
f = feats.cuda()
X = f[0].view(64, -1).t()
for cta in (0, 1, 73, 147):
    for variant in (0, 1, 2):
        os.environ["UOC_FPS_TRACE"] = str(cta)
        os.environ["UOC_FPS_VARIANT"] = str(variant)
        for rep in range(2):
            MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
        torch.cuda.synchronize()
        ws = MS._workspaces[("cuda", 0)]
        t = ws[: 3 * 100 * 8].view(torch.int64).cpu().numpy().reshape(100, 3)[:99]
        comp = (t[:, 1] - t[:, 0])
        sync = (t[:, 2] - t[:, 1])
        gap = (t[1:, 0] - t[:-1, 2])
        tot = (t[1:, 0] - t[:-1, 0])
        print("cta %3d variant %d: compute %.0f +- %.0f clk, exchange %.0f +- %.0f clk, gap %.0f, pass %.0f clk  (first passes compute %s exch %s)"
              % (cta, variant, comp[5:].mean(), comp[5:].std(), sync[5:].mean(), sync[5:].std(), gap[5:].mean(), tot[5:].mean(),
                 comp[:3].tolist(), sync[:3].tolist()))
