"""Host-side mirror of the reference's network factories for the ResNet34-8s embedding network
(lib/networks/SEG.py:173-181 `seg_resnet34_8s_embedding`, `seg_resnet34_8s_embedding_early`, class SEGNET :26-119).

The reference SEGNET reads cfg.INPUT / cfg.TRAIN.FUSION_TYPE / cfg.TRAIN.EMBEDDING_NORMALIZATION at
construction (SEG.py:34-38); the factories here read the same three settings from the module-level
CONFIG dict (defaults = the shipped rgbd_add configs; `configure(cfg)` copies them from a reference
cfg object), or take them as keyword arguments.  Variants: INPUT 'RGBD' | 'COLOR' | 'DEPTH';
FUSION_TYPE 'add' | 'cat' | 'early' (SEG.py:97-110).

The module keeps the weights as a reference-format state_dict (same keys, so reference
checkpoints load unchanged and .state_dict()/.cuda()/DataParallel behave), and runs the forward
pass through uoc_backbone_forward: BN-folded bf16 implicit-GEMM convolutions on tcgen05, fused
add + bilinear x8 + L2-normalise head.  Inference only.
"""
import collections
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from . import mean_shift as _ms

__all__ = ["seg_resnet34_8s_embedding", "seg_resnet34_8s_embedding_early", "SEGNET_B200", "reference_state_dict_keys",
           "random_state_dict", "CONFIG", "configure"]

_LAYERS = ((64, 3), (128, 4), (256, 6), (512, 3))

# what SEGNET.__init__ reads from the reference's global cfg (SEG.py:34-38); defaults = experiments/cfgs/*rgbd_add*.yml
CONFIG = {"INPUT": "RGBD", "FUSION_TYPE": "add", "EMBEDDING_NORMALIZATION": True}

_INPUT_IDS = {"RGBD": 0, "COLOR": 1, "DEPTH": 2}
_FUSION_IDS = {"add": 0, "cat": 1, "early": 2}


_LIVE_CFG = [None]      # a reference cfg object to read at construction time (set by shim.install)


def configure(cfg, live=False):
    """Copy INPUT / TRAIN.FUSION_TYPE / TRAIN.EMBEDDING_NORMALIZATION from a reference cfg (lib/fcn/config.py).
    live=True keeps the object instead and reads it whenever a network is constructed, like the reference's
    SEGNET.__init__ does (the tools call cfg_from_file() after importing the modules)."""
    if live:
        _LIVE_CFG[0] = cfg
        return
    CONFIG["INPUT"] = str(cfg.INPUT)
    CONFIG["FUSION_TYPE"] = str(cfg.TRAIN.FUSION_TYPE)
    CONFIG["EMBEDDING_NORMALIZATION"] = bool(cfg.TRAIN.EMBEDDING_NORMALIZATION)


def current_config():
    if _LIVE_CFG[0] is not None:
        configure(_LIVE_CFG[0])
    return CONFIG


def _branches(input_type, fusion_type, in_channels):
    """[(state-dict prefix, stem input channels)] of the trunks SEGNET builds (SEG.py:69-71)."""
    two = input_type == "RGBD" and fusion_type != "early"
    return [("fcn", in_channels)] + ([("fcn_depth", in_channels)] if two else [])


def reference_state_dict_keys(num_units=64, input_type="RGBD", fusion_type="add", in_channels=3):
    """(key, shape) for every tensor of the reference module's state_dict, in its order
    (lib/networks/resnet.py:141-186: conv1, bn1, layer1..4, fc; num_batches_tracked included)."""
    out = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                    (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())])

    for top, cin in _branches(input_type, fusion_type, in_channels):
        p = top + ".resnet34_8s."
        out.append((p + "conv1.weight", (64, cin, 7, 7)))
        bn(p + "bn1", 64)
        inplanes = 64
        for li, (planes, blocks) in enumerate(_LAYERS, start=1):
            for b in range(blocks):
                q = "%slayer%d.%d." % (p, li, b)
                out.append((q + "conv1.weight", (planes, inplanes, 3, 3)))
                bn(q + "bn1", planes)
                out.append((q + "conv2.weight", (planes, planes, 3, 3)))
                bn(q + "bn2", planes)
                if b == 0 and li > 1:
                    out.append((q + "downsample.0.weight", (planes, inplanes, 1, 1)))
                    bn(q + "downsample.1", planes)
                inplanes = planes
        out.append((p + "fc.weight", (num_units, 512, 1, 1)))
        out.append((p + "fc.bias", (num_units,)))
    return out


def random_state_dict(num_units=64, seed=None, input_type="RGBD", fusion_type="add", in_channels=3):
    """Random initialisation in the spirit of SEGNET._initialize_weights (SEG.py:77-85): xavier-normal
    convolutions, zero biases, BN weight 1 / bias 0 / running stats (0, 1)."""
    gen = torch.Generator()
    if seed is not None:
        gen.manual_seed(int(seed))
    else:
        gen.seed()
    sd = {}
    for k, shape in reference_state_dict_keys(num_units, input_type, fusion_type, in_channels):
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            fan_out = shape[0] * shape[2] * shape[3]
            std = math.sqrt(2.0 / (fan_in + fan_out))
            sd[k] = torch.randn(shape, generator=gen) * std
        elif k.endswith("running_var") or (k.endswith(".weight") and len(shape) == 1):
            sd[k] = torch.ones(shape)
        else:
            sd[k] = torch.zeros(shape)
    return sd


def _normalize_keys(data):
    """Key handling of SEG.py:130-159 and tools/test_net.py:111-112."""
    if isinstance(data, dict) and "model" in data and not torch.is_tensor(data["model"]):
        data = data["model"]
    out = {}
    for k, v in data.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


_MAX_GRAPHS = 16         # captured shapes kept per network (least recently used out)


class _GraphedForward(object):
    """One captured CUDA graph of uoc_backbone_forward with its static buffers (SEGNET_B200.forward_ex)."""

    def __init__(self, net, handle, dev, N, H, W):
        need_img, need_depth = net.input_type != "DEPTH", net.input_type != "COLOR"
        with torch.cuda.device(dev):
            nbytes = _lib.load().uoc_backbone_workspace_bytes(handle, N, H, W)
            self.ws = torch.empty(nbytes + 2048, dtype=torch.uint8, device=dev)
            self.img = torch.zeros((N, 3, H, W), dtype=torch.float32, device=dev) if need_img else None
            self.depth = torch.zeros((N, 3, H, W), dtype=torch.float32, device=dev) if need_depth else None
            self.out = torch.empty((N, net.feature_dim, H, W), dtype=torch.float32, device=dev)
            self.xb = torch.empty((N, H * W, net.feature_dim), dtype=torch.bfloat16, device=dev) if net.keep_bf16 else None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                       # once eagerly: lazy per-kernel attributes are set outside the capture
                net._launch(handle, dev, self.img, self.depth, N, H, W, self.out, self.xb, self.ws)
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                net._launch(handle, dev, self.img, self.depth, N, H, W, self.out, self.xb, self.ws)
            torch.cuda.current_stream(dev).wait_stream(side)


class SEGNET_B200(nn.Module):
    """Drop-in for the reference SEGNET built by seg_resnet34_8s_embedding[_early].  forward(img, label, depth)
    returns features [N, C, H, W] float32 on the input's device (C = num_units, or 2 * num_units for cat
    fusion; unit L2 norm over C when `normalize`)."""

    def __init__(self, num_units=64, data=None, flags=0, input_type=None, fusion_type=None, normalize=None,
                 in_channels=3):
        super().__init__()
        self.num_units = int(num_units)
        self.flags = int(flags)
        conf = current_config()
        self.input_type = str(conf["INPUT"] if input_type is None else input_type)
        self.fusion_type = str(conf["FUSION_TYPE"] if fusion_type is None else fusion_type)
        self.normalize = bool(conf["EMBEDDING_NORMALIZATION"] if normalize is None else normalize)
        self.in_channels = int(in_channels)
        if self.input_type not in _INPUT_IDS or self.fusion_type not in _FUSION_IDS:
            raise ValueError("INPUT must be RGBD / COLOR / DEPTH and FUSION_TYPE add / cat / early")
        early = self.input_type == "RGBD" and self.fusion_type == "early"
        if self.in_channels != (6 if early else 3):
            raise ValueError("in_channels must be 6 for early fusion (seg_resnet34_8s_embedding_early) and 3 otherwise")
        cat = self.input_type == "RGBD" and self.fusion_type == "cat"
        self.feature_dim = 2 * self.num_units if cat else self.num_units
        sd = random_state_dict(num_units, None, self.input_type, self.fusion_type, self.in_channels)
        if data is not None:
            given = _normalize_keys(data)
            for k, v in given.items():          # same filter as SEG.py:152: name and shape must match
                if k in sd and tuple(v.shape) == tuple(sd[k].shape):
                    sd[k] = v.detach().clone().to(sd[k].dtype)
        self._names = list(sd.keys())
        for k, v in sd.items():
            self.register_buffer(k.replace(".", "__"), v, persistent=True)
        self._handles = {}          # device -> uoc_backbone handle; shared (by reference) with DataParallel replicas
        self._owner = True          # replicas never destroy handles
        self._ws = None
        self._ws_by_stream = {}
        self._graphs = collections.OrderedDict()
        self._seen = {}
        self.use_graphs = True
        self.keep_bf16 = True
        self.eval()

    # reference-format state_dict -------------------------------------------------------------
    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return type(sd)((k.replace("__", "."), v) for k, v in sd.items())

    def load_state_dict(self, state_dict, strict=True, **kw):
        given = _normalize_keys(state_dict)
        mapped = {k.replace(".", "__"): v for k, v in given.items()}
        res = super().load_state_dict(mapped, strict=strict, **kw)
        self._release()
        return res

    def _release(self):
        self._graphs = collections.OrderedDict()      # captured graphs hold the handle's weights
        self._seen = {}
        if not getattr(self, "_owner", False):
            return
        handles, self._handles = self._handles, {}
        for h in handles.values():
            _lib.load().uoc_backbone_destroy(h)

    def _replicate_for_data_parallel(self):
        # torch.nn.DataParallel over several devices: replicas share the per-device handle table (one handle per
        # device, built once from the host copy of the weights) and own nothing themselves
        replica = super()._replicate_for_data_parallel()
        replica._owner = False
        replica._ws = None
        replica._ws_by_stream = {}
        replica._graphs = collections.OrderedDict()
        replica._seen = {}
        replica.use_graphs = False      # replicas live for one forward: nothing to amortise a capture over
        return replica

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):      # .cuda() / .to(): weights moved -> rebuild the handle lazily
        self._release()
        self._ws = None
        self._ws_by_stream = {}
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("SEGNET_B200 is inference only (the training path is out of scope)")
        return super().train(False)

    # handle ------------------------------------------------------------------------------------
    def _ensure_handle(self, device):
        h = self._handles.get(device)
        if h is not None:
            return h
        lib = _lib.load()
        keep = []
        descs = []
        for k in self._names:
            if k.endswith("num_batches_tracked"):
                continue
            t = getattr(self, k.replace(".", "__")).detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            descs.append(_lib.WeightDesc(k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()))
        arr = (_lib.WeightDesc * len(descs))(*descs)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.uoc_backbone_create_ex(ctypes.byref(h), arr, len(descs), self.num_units,
                                                  _INPUT_IDS[self.input_type], _FUSION_IDS[self.fusion_type],
                                                  int(self.normalize)), "uoc_backbone_create_ex")
        self._handles[device] = h
        return h

    def forward(self, img, label=None, depth=None):
        """features [N, C, H, W] float32 (the reference's return value).  The bf16 pixel-major copy written by the same
        kernel is remembered for exactly this tensor (mean_shift.register_bf16_copy) so that clustering_features(features)
        finds it; callers that own both use forward_ex."""
        out, xb = self.forward_ex(img, label, depth)
        if xb is not None:
            _ms.register_bf16_copy(out, xb)
        return out

    def forward_ex(self, img, label=None, depth=None, graph=None, static_outputs=False):
        """(features [N, C, H, W] float32, bf16 pixel-major copy [N, H*W, C] or None when keep_bf16 is off).

        graph (default: self.use_graphs): replay a captured CUDA graph of the ~45 launches of this (device, N, H, W)
        instead of enqueueing them one by one -- the host cost of a forward drops from ~1.7 ms to one replay.  The graph
        owns static input / output / workspace buffers: inputs are copied in, and the outputs are copied out into fresh
        tensors unless static_outputs=True (then they are the graph's own buffers, valid until the next forward of the
        same shape).  Inside somebody else's stream capture (pipeline.py) and under FLAG_SYNC_CHECK the launches are
        enqueued directly."""
        need_img, need_depth = self.input_type != "DEPTH", self.input_type != "COLOR"
        if need_depth and depth is None:
            raise _lib.UocError("INPUT=%r needs the depth (XYZ) tensor" % self.input_type)
        if need_img and img is None:
            raise _lib.UocError("INPUT=%r needs the image tensor" % self.input_type)
        lead = img if need_img else depth
        if not lead.is_cuda:
            raise _lib.UocError("inputs must be CUDA tensors: there is no CPU path in this package")
        dev = lead.device
        handle = self._ensure_handle(dev)
        N, _, H, W = lead.shape
        use_graph = self.use_graphs if graph is None else graph
        if use_graph and not (self.flags & _lib.FLAG_SYNC_CHECK) and not torch.cuda.is_current_stream_capturing():
            g = self._graph_for(handle, dev, N, H, W)
        else:
            g = None
        if g is not None:
            if need_img:
                g.img.copy_(img.detach(), non_blocking=True)
            if need_depth:
                g.depth.copy_(depth.detach(), non_blocking=True)
            g.graph.replay()
            self._ws = g.ws
            if static_outputs:
                return g.out, g.xb
            return g.out.clone(), (g.xb.clone() if g.xb is not None else None)
        img = img.detach().to(device=dev, dtype=torch.float32).contiguous() if need_img else None
        depth = depth.detach().to(device=dev, dtype=torch.float32).contiguous() if need_depth else None
        with torch.cuda.device(dev):
            sid = torch.cuda.current_stream(dev).cuda_stream       # one activation workspace per stream in flight
            self._ws = self._ws_by_stream.get(sid)
            nbytes = _lib.load().uoc_backbone_workspace_bytes(handle, N, H, W)
            if self._ws is None or self._ws.device != dev or self._ws.numel() < nbytes + 1024:
                self._ws = torch.empty(nbytes + 2048, dtype=torch.uint8, device=dev)
                self._ws_by_stream[sid] = self._ws
            out = torch.empty((N, self.feature_dim, H, W), dtype=torch.float32, device=dev)
            xb = torch.empty((N, H * W, self.feature_dim), dtype=torch.bfloat16, device=dev) if self.keep_bf16 else None
            self._launch(handle, dev, img, depth, N, H, W, out, xb, self._ws)
        return out, xb

    def _launch(self, handle, dev, img, depth, N, H, W, out, xb, ws):
        off = (-ws.data_ptr()) % 1024
        st = _lib.load().uoc_backbone_forward(handle, _lib.ptr(img), _lib.ptr(depth), N, H, W, _lib.ptr(out), _lib.ptr(xb),
                                              ctypes.c_void_p(ws.data_ptr() + off), ws.numel() - off, self.flags,
                                              _lib.stream_ptr(dev))
        _lib.check(st, "uoc_backbone_forward")

    def _graph_for(self, handle, dev, N, H, W):
        """The captured forward of one (device, shape, flags, knob state); at most _MAX_GRAPHS are kept (least recently used out)."""
        key = (dev, N, H, W, self.flags, self.keep_bf16, _lib.knob_epoch())
        g = self._graphs.pop(key, None)
        if g is None:
            # a shape is captured when it comes back (the crop network sees a different batch size per frame: capturing
            # every one-off shape would cost more than it saves)
            seen = self._seen.get(key, 0) + 1
            if len(self._seen) > 256:
                self._seen.clear()
            self._seen[key] = seen
            if seen < 2:
                return None
            g = _GraphedForward(self, handle, dev, N, H, W)
            while len(self._graphs) >= _MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = g                                      # most recently used last
        return g

    def read_trunk(self, branch, N, H, W):
        """Test hook: trunk output [N, num_units, H/8, W/8] of one branch from the last forward."""
        lib = _lib.load()
        dev = self._ws.device
        h3 = ((((H - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
        w3 = ((((W - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
        out = torch.empty((N, self.num_units, h3, w3), dtype=torch.float32, device=dev)
        off = (-self._ws.data_ptr()) % 1024
        with torch.cuda.device(dev):
            _lib.check(lib.uoc_backbone_read_trunk(self._handles[dev], branch, N, H, W,
                                                   ctypes.c_void_p(self._ws.data_ptr() + off), _lib.ptr(out),
                                                   _lib.stream_ptr(dev)), "uoc_backbone_read_trunk")
        return out


def seg_resnet34_8s_embedding(num_classes=2, num_units=64, data=None, **variant):
    """Same signature as lib/networks/SEG.py:173-176 (in_channels=3).  The variant (INPUT, FUSION_TYPE,
    EMBEDDING_NORMALIZATION) comes from CONFIG like the reference's from cfg, or from the keyword arguments
    input_type= / fusion_type= / normalize=."""
    return SEGNET_B200(num_units=num_units, data=data, in_channels=3, **variant)


def seg_resnet34_8s_embedding_early(num_classes=2, num_units=64, data=None, **variant):
    """Same signature as lib/networks/SEG.py:178-181: one trunk with a 6-channel stem on cat(img, depth)
    (meaningful with INPUT='RGBD', FUSION_TYPE='early', which is the default here)."""
    variant.setdefault("input_type", "RGBD")
    variant.setdefault("fusion_type", "early")
    return SEGNET_B200(num_units=num_units, data=data, in_channels=6, **variant)
