"""Per-layer device time of the backbone's convolution shapes through uoc_conv2d_bf16 (CUDA events, L2 flushed), for
A/B of kernel variants selected by library knobs (UOC_CONV_LAYERS_VARIANTS="auto;conv_pair=1;conv_pair=0,conv_trace=1").  N images emulate (frames per launch) x (2 branches).
usage: python tools/conv_layers.py [N ...]   (default 2 and 8 = batch 1 and batch 4 with both branches)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unseenobjectclustering_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# (name, H, W, Cin, Cout, k, stride, dil, count per trunk)
LAYERS = [("l1 3x3 64", 120, 160, 64, 64, 3, 1, 1, 6), ("l2 3x3 s2 64>128", 120, 160, 64, 128, 3, 2, 1, 1),
          ("l2 1x1 s2 down", 120, 160, 64, 128, 1, 2, 1, 1), ("l2 3x3 128", 60, 80, 128, 128, 3, 1, 1, 7),
          ("l3 3x3 128>256 d2", 60, 80, 128, 256, 3, 1, 2, 1), ("l3 1x1 down", 60, 80, 128, 256, 1, 1, 1, 1),
          ("l3 3x3 256 d2", 60, 80, 256, 256, 3, 1, 2, 11), ("l4 3x3 256>512 d4", 60, 80, 256, 512, 3, 1, 4, 1),
          ("l4 1x1 down", 60, 80, 256, 512, 1, 1, 1, 1), ("l4 3x3 512 d4", 60, 80, 512, 512, 3, 1, 4, 5),
          ("fc 1x1 512>64", 60, 80, 512, 64, 1, 1, 1, 1)]
VARIANTS = [("auto", {}), ("pair", {"conv_pair": "1", "conv_wres": "0"}), ("tc", {"conv_pair": "0", "conv_wres": "0"})]
CHAIN = int(os.environ.get("UOC_CONV_LAYERS_CHAIN", "1"))   # > 1: that many launches back to back per measurement (warm L2, no launch gaps)
if os.environ.get("UOC_CONV_LAYERS_VARIANTS"):
    VARIANTS = [(v, dict(kv.split("=") for kv in v.split(",") if "=" in kv)) for v in os.environ["UOC_CONV_LAYERS_VARIANTS"].split(";")]
if os.environ.get("UOC_CONV_LAYERS_ONLY"):
    keep = os.environ["UOC_CONV_LAYERS_ONLY"].split(",")
    LAYERS = [l for l in LAYERS if l[0] in keep]
out = {}
for N in ([int(v) for v in sys.argv[1:]] or [2, 8]):
    for name, H, W, Cin, Cout, k, stride, dil, cnt in LAYERS:
        x = torch.randn(N, H, W, Cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(Cout, k * k, Cin, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.zeros(Cout, device=dev)
        pad = dil if k == 3 else 0
        Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
        y = torch.empty(N, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        row = {}
        for vname, env in VARIANTS:
            for kk in ("conv_pair", "conv_wres", "conv_debug", "conv_trace"):        # library knobs (uoc_set_knob), back to their defaults
                _lib.set_knob(kk, _lib.KNOB_DEFAULTS[kk])
            for kk, vv in env.items():
                _lib.set_knob(kk, int(vv))
            ts = []
            for rep in range(7):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(CHAIN):
                    _lib.check(lib.uoc_conv2d_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), None, _lib.ptr(y), N, H, W, Cin, Cout, k, stride,
                                                   dil, 1, 0, _lib.stream_ptr(dev)), "conv")
                e.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    ts.append(s.elapsed_time(e) * 1e3 / CHAIN)
            us = sorted(ts)[len(ts) // 2]
            row[vname] = {"us": round(us, 2), "TFLOPs": round(flops / us / 1e6, 1)}
        out["N%d %s" % (N, name)] = dict(row, count=cnt, gflop=round(flops / 1e9, 2))
        print("N=%d %-20s x%-2d %7.2f GF " % (N, name, cnt, flops / 1e9) + "  ".join("%s %7.1f us %6.0f TF/s" % (v, r["us"], r["TFLOPs"]) for v, r in row.items()), flush=True)
    for vname, _ in VARIANTS:
        tot = sum(v[vname]["us"] * v["count"] for kk, v in out.items() if kk.startswith("N%d " % N))
        print("N=%d sum over the trunk's convolutions (%s): %.1f us" % (N, vname, tot), flush=True)
json.dump(out, open("gpurun_out/conv_layers%s.json" % os.environ.get("UOC_AB_TAG", ""), "w"), indent=1)
