"""Host-side mirror of the reference's lib/utils/mean_shift.py on top of the CUDA C ABI.

Same function names, argument meaning and return types as the reference; the arithmetic runs in
libuoc_b200.so (farthest point sampling, tcgen05 mean-shift loop, greedy seed labelling, nearest-seed
assignment).  metric='cosine' (every shipped config: experiments/cfgs/*.yml `EMBEDDING_METRIC: cosine`)
takes the tensor-core kernels; metric='euclidean' (the cfg default, lib/fcn/config.py:261) runs the same
stages on the fp32 SIMT kernels.  PyTorch is used for device memory and streams only.
"""
import ctypes

import numpy as np
import torch

from . import _lib

EMBEDDING_ALPHA = 0.02   # cfg.TRAIN.EMBEDDING_ALPHA default, lib/fcn/config.py:254

_workspaces = {}
_bf16_registry = {}      # data_ptr -> (the feature tensor itself, its bf16 pixel-major copy, its version) made by the backbone
_BF16_REGISTRY_SIZE = 2      # feature maps kept alive (a frame and its crops); FramePipeline and bench own their copies explicitly


def _workspace(device, nbytes):
    """Grow-only uint8 scratch tensor per (device, current stream): frames in flight on different
    streams (pipeline.py) must not share scratch.  256-byte aligned by the caching allocator."""
    sid = torch.cuda.current_stream(device).cuda_stream
    key = (device.type, device.index) if sid == 0 else (device.type, device.index, sid)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def register_bf16_copy(features, xb):
    """Called by the backbone module: remember the bf16 [N, H*W, C] copy written next to `features`.

    The entry holds a STRONG reference to `features`, so its address cannot be handed to another tensor by the
    caching allocator while the entry is live (an entry keyed by a bare data_ptr could silently alias a later,
    unrelated field of the same shape).  At most _BF16_REGISTRY_SIZE entries (feature maps) are kept alive."""
    key = features.data_ptr()
    _bf16_registry.pop(key, None)
    while len(_bf16_registry) >= _BF16_REGISTRY_SIZE:
        _bf16_registry.pop(next(iter(_bf16_registry)))
    _bf16_registry[key] = (features, xb, features._version)


def clear_bf16_registry():
    """Drop every remembered (features, bf16 copy) pair (releases the feature maps the registry keeps alive)."""
    _bf16_registry.clear()


def _lookup_bf16(features):
    """The bf16 copy the backbone wrote next to `features`, or None.  A hit requires the SAME storage, offset, shape and
    strides as the registered tensor (the tensor itself or an alias such as .detach()) and an unchanged autograd version
    counter (an in-place torch op on the field since the forward pass invalidates the copy)."""
    ent = _bf16_registry.get(features.data_ptr())
    if ent is None:
        return None
    owner, xb, version = ent
    same = (features is owner or
            (features.untyped_storage().data_ptr() == owner.untyped_storage().data_ptr()
             and features.storage_offset() == owner.storage_offset()))
    if (same and features.dtype == owner.dtype and tuple(features.shape) == tuple(owner.shape)
            and tuple(features.stride()) == tuple(owner.stride())
            and features._version == version and owner._version == version):
        return xb
    if features is owner or same:
        _bf16_registry.pop(features.data_ptr(), None)      # modified in place: the copy is stale for good
    return None


def _epsilon(epsilon=None):
    return float(2 * EMBEDDING_ALPHA) if epsilon is None else float(epsilon)   # mean_shift.py:123


def _metric_flag(metric):
    if metric == 'cosine':
        return 0
    if metric == 'euclidean':
        return _lib.FLAG_EUCLIDEAN
    raise ValueError("metric must be 'cosine' or 'euclidean' (cfg.TRAIN.EMBEDDING_METRIC)")


def cluster_fields(features, num_seeds=100, kappa=20.0, max_iters=10, first_indices=None, epsilon=None, flags=0,
                   return_seeds=False, on_sampling_done=None, metric='cosine', x_bf16=None, labels_f32_out=None,
                   labels_u8_out=None):
    """Cluster a batch of embedding fields in ONE library call.

    features: [N, C, H, W] float32 CUDA tensor, unit norm over C (any batch stride; each item must
    be planar-contiguous like the reference's network output).
    first_indices: the per-item first seed (np.random.randint(0, n) of mean_shift.py:155); drawn
    here from numpy's global RNG, in item order, when None.
    x_bf16: the bf16 pixel-major copy [N, H*W, C] of `features` when the caller owns it (SEGNET_B200.forward_ex);
    None -> the copy the backbone registered for exactly this tensor, else it is made inside the library.
    labels_f32_out / labels_u8_out: optional contiguous CUDA tensors with N * H*W elements (float32 / uint8) that receive
    the label map in that type from the same final pass (the reference's out_label is float32; uint8 is the gather type).
    Returns (labels int32 [N, H*W] CUDA, selected int64 [N, num_seeds] CUDA[, seeds, seed_labels]).
    """
    if not features.is_cuda:
        raise _lib.UocError("features must be a CUDA tensor: there is no CPU path in this package")
    if features.dtype != torch.float32:
        raise _lib.UocError("features must be float32")
    N, C, H, W = features.shape
    n = H * W
    xb = None
    if metric == 'cosine':
        xb = x_bf16 if x_bf16 is not None else _lookup_bf16(features)
        if xb is not None and (xb.dtype != torch.bfloat16 or tuple(xb.shape) != (N, n, C) or not xb.is_contiguous()
                               or xb.device != features.device):
            raise _lib.UocError("x_bf16 must be a contiguous bfloat16 [N, H*W, C] tensor on the features' device")
    if not (features.stride(3) == 1 and features.stride(2) == W and features.stride(1) == n):
        features = features.contiguous()
    lib = _lib.load()
    dev = features.device
    if first_indices is None:
        first_indices = [np.random.randint(0, n) for _ in range(N)]
    if len(first_indices) < N:
        raise ValueError("first_indices has %d entries for %d fields" % (len(first_indices), N))
    for t, dt in ((labels_f32_out, torch.float32), (labels_u8_out, torch.uint8)):
        if t is not None and (t.dtype != dt or t.numel() != N * n or not t.is_contiguous() or t.device != features.device):
            raise _lib.UocError("labels_f32_out / labels_u8_out must be contiguous CUDA tensors with N*H*W elements")
    first = (ctypes.c_int64 * N)(*[int(v) for v in list(first_indices)[:N]])     # extra entries are left unused
    flags = int(flags) | _metric_flag(metric)
    with torch.cuda.device(dev):
        nbytes = lib.uoc_meanshift_workspace_bytes(N, n, C, num_seeds)
        ws = _workspace(dev, nbytes)
        labels = torch.empty((N, n), dtype=torch.int32, device=dev)
        selected = torch.empty((N, num_seeds), dtype=torch.int64, device=dev)
        if on_sampling_done is not None:
            # staged form (same kernels through the stage entry points): lets the caller record an event right after
            # the cooperative sampling kernel so that the next frame's sampling may start (pipeline.py)
            sp = _lib.stream_ptr(dev)
            fptr = ctypes.cast(first, ctypes.c_void_p)
            seeds = torch.empty((N, num_seeds, C), dtype=torch.float32, device=dev)
            seed_labels = torch.empty((N, num_seeds), dtype=torch.int32, device=dev)
            nuniq = torch.empty((N,), dtype=torch.int32, device=dev)
            sb, sd_ = features.stride(0), features.stride(1)
            _lib.check(lib.uoc_select_seeds(_lib.ptr(features), sb, sd_, _lib.ptr(xb), N, n, C, num_seeds, fptr,
                                            _lib.ptr(selected), _lib.ptr(seeds), _lib.ptr(ws), ws.numel(),
                                            int(flags) & (_lib.FLAG_FPS_FP32 | _lib.FLAG_EUCLIDEAN), sp),
                       "uoc_select_seeds")
            on_sampling_done()
            _lib.check(lib.uoc_hill_climb(_lib.ptr(features), sb, sd_, _lib.ptr(xb), N, n, C, num_seeds, float(kappa),
                                          int(max_iters), _lib.ptr(seeds), _lib.ptr(ws), ws.numel(), int(flags), sp),
                       "uoc_hill_climb")
            _lib.check(lib.uoc_label_seeds_ex(_lib.ptr(seeds), N, num_seeds, C, _epsilon(epsilon), int(flags),
                                              _lib.ptr(seed_labels), _lib.ptr(nuniq), sp), "uoc_label_seeds")
            _lib.check(lib.uoc_assign_labels_typed(_lib.ptr(features), sb, sd_, _lib.ptr(xb), N, n, C, num_seeds, _lib.ptr(seeds),
                                                   _lib.ptr(seed_labels), _lib.ptr(nuniq), _lib.ptr(labels),
                                                   _lib.ptr(labels_f32_out), _lib.ptr(labels_u8_out), _lib.ptr(ws),
                                                   ws.numel(), int(flags), sp), "uoc_assign_labels")
            if return_seeds:
                return labels, selected, seeds, seed_labels
            return labels, selected
        seeds = torch.empty((N, num_seeds, C), dtype=torch.float32, device=dev) if return_seeds else None
        seed_labels = torch.empty((N, num_seeds), dtype=torch.int32, device=dev) if return_seeds else None
        st = lib.uoc_meanshift_cluster_ex(
            _lib.ptr(features), features.stride(0), features.stride(1), _lib.ptr(xb), N, n, C, num_seeds,
            float(kappa), int(max_iters), _epsilon(epsilon), ctypes.cast(first, ctypes.c_void_p), _lib.ptr(labels),
            _lib.ptr(labels_f32_out), _lib.ptr(labels_u8_out), _lib.ptr(selected), _lib.ptr(seeds), _lib.ptr(seed_labels),
            _lib.ptr(ws), ws.numel(), int(flags), _lib.stream_ptr(dev))
        _lib.check(st, "uoc_meanshift_cluster")
    if return_seeds:
        return labels, selected, seeds, seed_labels
    return labels, selected


def _as_planar(X):
    """[n, d] tensor -> (tensor, stride_d) with points contiguous (the reference passes the
    transposed view of an NCHW feature map, strides (1, n): lib/fcn/test_dataset.py:54-55)."""
    if X.stride(0) == 1 and X.stride(1) >= X.shape[0]:
        return X, X.stride(1)
    Xp = X.t().contiguous()          # [d, n]
    return Xp.t(), Xp.stride(0)


def mean_shift_smart_init(X, kappa, num_seeds=100, max_iters=10, metric='cosine', first_index=None, flags=0):
    """lib/utils/mean_shift.py:192-229.  X: [n, d] CUDA float32 unit rows.  Returns
    (cluster_labels int64 [n] on the CPU, selected_indices int64 [num_seeds] on the CPU), like the
    reference."""
    n, d = X.shape
    Xp, stride_d = _as_planar(X)
    feats = torch.as_strided(Xp, (1, d, 1, n), (d * stride_d, stride_d, n, 1))
    fi = None if first_index is None else [first_index]
    labels, selected = cluster_fields(feats, num_seeds, kappa, max_iters, fi, flags=flags, metric=metric)
    return labels[0].to(torch.int64).cpu(), selected[0].cpu()


def select_smart_seeds(X, num_seeds, return_selected_indices=False, init_seeds=None, num_init_seeds=None,
                       metric='cosine', first_index=None, bf16_screen=True):
    """lib/utils/mean_shift.py:128-189.  init_seeds [num_seeds, d] with its first num_init_seeds rows chosen already
    (:144-149): sampling continues from them; like the reference, the remaining rows of init_seeds are filled IN PLACE, the
    same tensor is returned, and selected_indices is -1 for the given rows.
    bf16_screen: screen every pass with a bf16 copy of X (d = 64/128; same indices, see fps_tc.cu; fresh starts only)."""
    mflag = _metric_flag(metric)
    n, d = X.shape
    Xp, stride_d = _as_planar(X)
    lib = _lib.load()
    dev = X.device
    if init_seeds is not None and num_init_seeds:
        k = int(num_init_seeds)
        given = init_seeds[:k].detach().to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            ws = _workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, d, num_seeds))
            selected = torch.empty((num_seeds,), dtype=torch.int64, device=dev)
            seeds = torch.empty((num_seeds, d), dtype=torch.float32, device=dev)
            st = lib.uoc_select_seeds_init(_lib.ptr(Xp), d * stride_d, stride_d, 1, n, d, num_seeds, _lib.ptr(given), k,
                                           _lib.ptr(selected), _lib.ptr(seeds), _lib.ptr(ws), ws.numel(),
                                           _lib.FLAG_SYNC_CHECK | mflag, _lib.stream_ptr(dev))
            _lib.check(st, "uoc_select_seeds_init")
        init_seeds[k:num_seeds] = seeds[k:].to(init_seeds.device, init_seeds.dtype)      # seeds = init_seeds (:147)
        if return_selected_indices:
            return init_seeds, selected.cpu()
        return (init_seeds,)
    if first_index is None:
        first_index = np.random.randint(0, n)
    first = (ctypes.c_int64 * 1)(int(first_index))
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, d, num_seeds))
        selected = torch.empty((num_seeds,), dtype=torch.int64, device=dev)
        seeds = torch.empty((num_seeds, d), dtype=torch.float32, device=dev)
        xb = None
        if bf16_screen and d in (64, 128) and not mflag:
            xb = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
            _lib.check(lib.uoc_pack_bf16(_lib.ptr(Xp), d * stride_d, stride_d, 1, n, d, _lib.ptr(xb), _lib.stream_ptr(dev)),
                       "uoc_pack_bf16")
        st = lib.uoc_select_seeds(_lib.ptr(Xp), d * stride_d, stride_d, _lib.ptr(xb), 1, n, d, num_seeds,
                                  ctypes.cast(first, ctypes.c_void_p), _lib.ptr(selected), _lib.ptr(seeds), _lib.ptr(ws),
                                  ws.numel(), _lib.FLAG_SYNC_CHECK | mflag, _lib.stream_ptr(dev))
        _lib.check(st, "uoc_select_seeds")
    if return_selected_indices:
        return seeds, selected.cpu()
    return (seeds,)


def seed_hill_climbing_ball(X, Z, kappa, max_iters=10, metric='cosine', flags=0):
    """lib/utils/mean_shift.py:79-109.  Returns the updated seeds (new tensor)."""
    flags = int(flags) | _metric_flag(metric)
    n, d = X.shape
    m = Z.shape[0]
    Xp, stride_d = _as_planar(X)
    lib = _lib.load()
    dev = X.device
    Zc = Z.detach().to(torch.float32).contiguous().clone()
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, d, m))
        st = lib.uoc_hill_climb(_lib.ptr(Xp), d * stride_d, stride_d, None, 1, n, d, m, float(kappa), int(max_iters),
                                _lib.ptr(Zc), _lib.ptr(ws), ws.numel(), int(flags) | _lib.FLAG_SYNC_CHECK,
                                _lib.stream_ptr(dev))
        _lib.check(st, "uoc_hill_climb")
    return Zc


def connected_components(Z, epsilon, metric='cosine', return_num_unique=False):
    """lib/utils/mean_shift.py:41-76.  Z: [m, d] CUDA float32.  Returns int64 labels on the CPU."""
    mflag = _metric_flag(metric)
    m, d = Z.shape
    lib = _lib.load()
    dev = Z.device
    Zc = Z.detach().to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        labels = torch.empty((m,), dtype=torch.int32, device=dev)
        num = torch.empty((1,), dtype=torch.int32, device=dev)
        st = lib.uoc_label_seeds_ex(_lib.ptr(Zc), 1, m, d, float(epsilon), mflag, _lib.ptr(labels), _lib.ptr(num),
                                    _lib.stream_ptr(dev))
        _lib.check(st, "uoc_label_seeds")
    if return_num_unique:
        return labels.to(torch.int64).cpu(), int(num.item())
    return labels.to(torch.int64).cpu()


def assign_labels(X, Z, seed_labels, num_unique=None, use_tensor_cores=False, metric='cosine'):
    """lib/utils/mean_shift.py:206-227: nearest-seed labels with the label-0 swap. int64 CPU [n].
    use_tensor_cores: certified tcgen05 pass + fp32 fix-up instead of the fp32 kernel (identical labels)."""
    n, d = X.shape
    m = Z.shape[0]
    Xp, stride_d = _as_planar(X)
    lib = _lib.load()
    dev = X.device
    sl = seed_labels.to(device=dev, dtype=torch.int32).contiguous()
    if num_unique is None:
        num_unique = int(torch.unique(sl).numel())
    nu = torch.tensor([num_unique], dtype=torch.int32, device=dev)
    Zc = Z.detach().to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, d, m))
        out = torch.empty((n,), dtype=torch.int32, device=dev)
        xb = None
        mflag = _metric_flag(metric)
        if use_tensor_cores and not mflag:
            xb = pack_bf16(torch.as_strided(Xp, (1, d, 1, n), (d * stride_d, stride_d, n, 1)).contiguous().view(1, d, 1, n))
        st = lib.uoc_assign_labels_ex(_lib.ptr(Xp), d * stride_d, stride_d, _lib.ptr(xb), 1, n, d, m, _lib.ptr(Zc), _lib.ptr(sl),
                                      _lib.ptr(nu), _lib.ptr(out), _lib.ptr(ws), ws.numel(), mflag, _lib.stream_ptr(dev))
        _lib.check(st, "uoc_assign_labels")
    return out.to(torch.int64).cpu()


def pack_bf16(features):
    """[N, C, H, W] fp32 planar -> [N, H*W, C] bf16 pixel-major (the layout the tcgen05 loop streams)."""
    N, C, H, W = features.shape
    features = features.contiguous()
    lib = _lib.load()
    dev = features.device
    out = torch.empty((N, H * W, C), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        st = lib.uoc_pack_bf16(_lib.ptr(features), C * H * W, H * W, N, H * W, C, _lib.ptr(out), _lib.stream_ptr(dev))
        _lib.check(st, "uoc_pack_bf16")
    return out
