"""A/B of the persistent mean-shift loop kernel's MUFU / FMA-pipe split (UOC_LOOP_POLY, read at every launch):
time per launch (CUDA events, L2 flushed between launches) and the effect on the converged seeds."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import _lib, mean_shift as MS, synthetic

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for (H, W, D, K, iters) in ((480, 640, 64, 6, 10), (720, 960, 128, 12, 30)):
    feats, _ = synthetic.clustered_features(H, W, D, K, 0.05, seed=0)
    feats = feats.to(dev)
    n, M = H * W, 100
    xb = MS.pack_bf16(feats)
    ws = MS._workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, D, M))
    sel = torch.empty((1, M), dtype=torch.int64, device=dev)
    Z0 = torch.empty((1, M, D), dtype=torch.float32, device=dev)
    first = (ctypes.c_int64 * 1)(n // 3)
    sp = _lib.stream_ptr(dev)
    _lib.check(lib.uoc_select_seeds(_lib.ptr(feats), D * n, n, _lib.ptr(xb), 1, n, D, M, ctypes.cast(first, ctypes.c_void_p),
                                    _lib.ptr(sel), _lib.ptr(Z0), _lib.ptr(ws), ws.numel(), 0, sp), "select_seeds")
    ref = None
    for poly in (0, 8, 12, 0, 8, 12):
        os.environ["UOC_LOOP_POLY"] = str(poly)
        times = []
        for rep in range(12):
            Z = Z0.clone()
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.uoc_hill_climb(_lib.ptr(feats), D * n, n, _lib.ptr(xb), 1, n, D, M, 20.0, iters, _lib.ptr(Z),
                                          _lib.ptr(ws), ws.numel(), _lib.FLAG_SYNC_CHECK, sp), "hill_climb")
            b.record()
            torch.cuda.synchronize()
            if rep >= 2:
                times.append(a.elapsed_time(b))
        times.sort()
        if ref is None:
            ref = Z.clone()
        cosd = float((1.0 - (Z * ref).sum(-1)).abs().max())
        key = "%dx%dx%d_T%d_poly%d" % (H, W, D, iters, poly)
        out.setdefault(key, []).append({"ms_median": times[len(times) // 2], "ms_min": times[0], "max_cos_dist_vs_poly0": cosd})
        print(key, out[key][-1], flush=True)
json.dump(out, open(os.path.join("gpurun_out", "loop_ab.json"), "w"), indent=1)
