"""Throughput of the frame path when B frames go through every kernel together (backbone N = B, clustering batch = B):
the latency-bound cooperative kernels (sampling, mean-shift loop) share their exchange steps between the frames."""
import json, os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import _lib, mean_shift as MS, networks, synthetic

dev = torch.device("cuda:0")
lib = _lib.load()
H, W, D, M = 480, 640, 64, 100
net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for B in ([int(v) for v in sys.argv[1:]] or [1, 2, 3, 4]):
    img, xyz = synthetic.rgbd_frame(H, W, seed=0, batch=B)
    img, xyz = img.to(dev), xyz.to(dev)
    firsts = [1000 + 17 * i for i in range(B)]
    n = H * W
    ws = MS._workspace(dev, lib.uoc_meanshift_workspace_bytes(B, n, D, M))
    sel = torch.empty((B, M), dtype=torch.int64, device=dev)
    Z = torch.empty((B, M, D), dtype=torch.float32, device=dev)
    sl = torch.empty((B, M), dtype=torch.int32, device=dev)
    nu = torch.empty((B,), dtype=torch.int32, device=dev)
    lab = torch.empty((B, n), dtype=torch.int32, device=dev)
    sp = _lib.stream_ptr(dev)
    first = (ctypes.c_int64 * B)(*firsts)
    acc = [0.0] * 4
    reps = 8
    for rep in range(reps + 2):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda._sleep(6_000_000)
        ev[0].record()
        feats = net(img, None, xyz)
        xb = MS._lookup_bf16(feats)
        ev[1].record()
        _lib.check(lib.uoc_select_seeds(_lib.ptr(feats), D * n, n, _lib.ptr(xb), B, n, D, M, ctypes.cast(first, ctypes.c_void_p),
                                        _lib.ptr(sel), _lib.ptr(Z), _lib.ptr(ws), ws.numel(), 0, sp), "select_seeds")
        ev[2].record()
        _lib.check(lib.uoc_hill_climb(_lib.ptr(feats), D * n, n, _lib.ptr(xb), B, n, D, M, 20.0, 10, _lib.ptr(Z),
                                      _lib.ptr(ws), ws.numel(), 0, sp), "hill_climb")
        ev[3].record()
        _lib.check(lib.uoc_label_seeds(_lib.ptr(Z), B, M, D, 0.04, _lib.ptr(sl), _lib.ptr(nu), sp), "label_seeds")
        _lib.check(lib.uoc_assign_labels(_lib.ptr(feats), D * n, n, _lib.ptr(xb), B, n, D, M, _lib.ptr(Z), _lib.ptr(sl), _lib.ptr(nu),
                                         _lib.ptr(lab), _lib.ptr(ws), ws.numel(), sp), "assign_labels")
        ev[4].record()
        torch.cuda.synchronize()
        if rep >= 2:
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) / reps
    tot = sum(acc)
    out[B] = {"backbone_ms": acc[0], "fps_ms": acc[1], "loop_ms": acc[2], "labels_ms": acc[3], "total_ms": tot,
              "ms_per_frame": tot / B, "frames_per_s": 1000.0 * B / tot}
    print(B, {k: round(v, 4) for k, v in out[B].items()}, flush=True)
json.dump(out, open("gpurun_out/batch_ab%s.json" % os.environ.get("UOC_AB_TAG", ""), "w"), indent=1)
