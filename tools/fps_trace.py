"""Debug: per-pass phase timing of the sampling kernel (UOC_FPS_TRACE=<cta>) on a 640x480x64 field."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectclustering_b200 import mean_shift as MS, synthetic

feats, _ = synthetic.clustered_features(480, 640, 64, 6, 0.05, 0)
f = feats.cuda()
X = f[0].view(64, -1).t()
for cta in (0,):
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

# all CTAs at pass 50
for variant in (2,):
    os.environ["UOC_FPS_TRACE"] = "-1"
    os.environ["UOC_FPS_VARIANT"] = str(variant)
    for rep in range(2):
        MS.select_smart_seeds(X, 100, return_selected_indices=True, first_index=71530)
    torch.cuda.synchronize()
    ws = MS._workspaces[("cuda", 0)]
    t = ws[: 8 * 148 * 8].view(torch.int64).cpu().numpy().reshape(148, 8)
    comp = t[:, 1] - t[:, 0]
    exch = t[:, 2] - t[:, 1]
    print("   argmax->stores issued: median %d ; stores->poll done: median %d (min %d max %d) ; poll done->seed ready: median %d (min %d max %d)" % (
        np.median(t[:, 4] - t[:, 1]), np.median(t[:, 5] - t[:, 4]), (t[:, 5] - t[:, 4]).min(), (t[:, 5] - t[:, 4]).max(),
        np.median(t[:, 2] - t[:, 5]), (t[:, 2] - t[:, 5]).min(), (t[:, 2] - t[:, 5]).max()))
    order = np.argsort(-comp)
    print("pass 50, variant %d: compute min %d median %d max %d ; exchange min %d median %d max %d" % (
        variant, comp.min(), np.median(comp), comp.max(), exch.min(), np.median(exch), exch.max()))
    print("slowest CTAs (cta, smid, compute, exchange):", [(int(c), int(t[c, 3]), int(comp[c]), int(exch[c])) for c in order[:12]])
    print("fastest CTAs:", [(int(c), int(t[c, 3]), int(comp[c]), int(exch[c])) for c in order[-6:]])
    # clocks are per-SM and not synchronised; durations are what matter
