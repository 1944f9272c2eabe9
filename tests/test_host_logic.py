"""Host-side plumbing of the product package checked against the oracle on CPU tensors
(these functions are device agnostic tensor plumbing; the CUDA kernels are covered by -m gpu)."""
import numpy as np
import pytest
import torch

import uoc_oracle as O
from unseenobjectclustering_b200 import networks as N
from unseenobjectclustering_b200 import test_dataset as TD


def _labels(H, W, K, seed):
    _, gt = O.synthetic_clustered_features(H, W, 8, K, 0.05, seed)
    return gt


def test_two_stage_plumbing_has_no_cpu_path():
    """filter_labels_depth / crop_rois / match_label_crop run in csrc/refine.cu only: CPU pixel tensors are refused
    (the kernels themselves are checked against the oracle by tests/test_gpu_pipeline.py)."""
    from unseenobjectclustering_b200 import _lib
    H, W = 32, 48
    lab = _labels(H, W, 3, 0).float()[None]
    img, xyz = O.synthetic_rgbd_frame(H, W, 0)
    with pytest.raises(_lib.UocError):
        TD.filter_labels_depth(lab, xyz, 0.8)
    with pytest.raises(_lib.UocError):
        TD.crop_rois(img, lab, xyz)
    with pytest.raises(_lib.UocError):
        TD.match_label_crop(lab, torch.zeros(1, 224, 224), torch.zeros(1, 224, 224), torch.zeros(1, 4), None)
    src = open(TD.__file__).read()
    assert "torch.nn.functional" not in src and "F.interpolate" not in src


def test_depth_is_gated_on_the_network_input_type():
    """lib/fcn/test_dataset.py:236-239: depth is used iff cfg.INPUT is DEPTH / RGBD."""
    class _Net(object):
        def __init__(self, it):
            self.input_type = it
    assert TD._uses_depth(_Net("RGBD")) and TD._uses_depth(_Net("DEPTH")) and not TD._uses_depth(_Net("COLOR"))
    wrapped = type("DP", (), {"module": _Net("COLOR")})()
    assert not TD._uses_depth(wrapped)
    assert TD._uses_depth(lambda *a: None)           # foreign callable, no live cfg: the reference default (RGBD configs)


def test_bf16_registry_cannot_alias_a_foreign_tensor():
    """The side channel from the backbone to clustering_features is keyed by identity, storage and version."""
    from unseenobjectclustering_b200 import mean_shift as MS
    MS.clear_bf16_registry()
    f = torch.zeros(1, 4, 2, 3)
    xb = torch.zeros(1, 6, 4, dtype=torch.bfloat16)
    MS.register_bf16_copy(f, xb)
    assert MS._lookup_bf16(f) is xb
    assert MS._lookup_bf16(f.detach()) is xb                       # an alias of the same storage (test_dataset.py:247)
    assert MS._lookup_bf16(f[:, :2]) is None                       # a view with another shape
    assert MS._lookup_bf16(torch.zeros(1, 4, 2, 3)) is None        # a foreign tensor
    f.add_(1.0)                                                    # modified in place: the copy is stale
    assert MS._lookup_bf16(f) is None and MS._lookup_bf16(f) is None
    # the registry keeps the registered tensor alive, so its address cannot be recycled while the entry is live
    g = torch.zeros(1, 4, 2, 3)
    addr = g.data_ptr()
    MS.register_bf16_copy(g, xb)
    del g
    others = [torch.zeros(1, 4, 2, 3) for _ in range(64)]
    assert all(o.data_ptr() != addr for o in others)
    assert all(MS._lookup_bf16(o) is None for o in others)
    for k in range(MS._BF16_REGISTRY_SIZE + 3):                    # bounded
        MS.register_bf16_copy(torch.zeros(1, 4, 2, 3), xb)
    assert len(MS._bf16_registry) == MS._BF16_REGISTRY_SIZE
    MS.clear_bf16_registry()


def test_module_state_dict_roundtrip_and_key_filter():
    sd = O.randomise_bn_(N.random_state_dict(64, seed=1), 2)
    net = N.seg_resnet34_8s_embedding(2, 64, sd)
    out = net.state_dict()
    assert list(out.keys()) == [k for k, _ in N.reference_state_dict_keys(64)]
    for k in sd:
        assert torch.equal(out[k], sd[k]), k
    # 'module.' prefix, {'model': ...} wrapper, wrong shapes silently skipped (SEG.py:145-152)
    wrapped = {"model": {"module." + k: v for k, v in sd.items()}}
    wrapped["model"]["module.fcn.resnet34_8s.fc.weight"] = torch.zeros(3, 3)
    net2 = N.seg_resnet34_8s_embedding(2, 64, wrapped)
    out2 = net2.state_dict()
    assert torch.equal(out2["fcn.resnet34_8s.conv1.weight"], sd["fcn.resnet34_8s.conv1.weight"])
    assert out2["fcn.resnet34_8s.fc.weight"].shape == (64, 512, 1, 1)
    assert not net.training
    with pytest.raises(NotImplementedError):
        net.train()
    dp = torch.nn.DataParallel(net2, device_ids=None) if False else net2   # constructing DataParallel needs a GPU
    assert dp is net2


def test_inputs_must_be_cuda():
    from unseenobjectclustering_b200 import _lib
    net = N.seg_resnet34_8s_embedding(2, 64, None)
    with pytest.raises(_lib.UocError):
        net(torch.zeros(1, 3, 32, 32), None, torch.zeros(1, 3, 32, 32))


def test_variant_configuration_and_factories():
    """The factories take INPUT / FUSION_TYPE / EMBEDDING_NORMALIZATION like the reference's SEGNET.__init__ takes them
    from cfg (SEG.py:34-38): module-level CONFIG, configure(cfg), keyword arguments; construction needs no GPU."""
    net = N.seg_resnet34_8s_embedding(2, 64, None)
    assert (net.input_type, net.fusion_type, net.normalize, net.feature_dim) == ("RGBD", "add", True, 64)
    assert len(net.state_dict()) == 436                      # 436 tensors for rgbd_add
    cat = N.seg_resnet34_8s_embedding(2, 64, None, fusion_type="cat")
    assert cat.feature_dim == 128
    color = N.seg_resnet34_8s_embedding(2, 64, None, input_type="COLOR")
    assert len(color.state_dict()) == 218 and all(k.startswith("fcn.") for k in color.state_dict())
    early = N.seg_resnet34_8s_embedding_early(2, 64, None)
    assert early.state_dict()["fcn.resnet34_8s.conv1.weight"].shape == (64, 6, 7, 7)

    class _T(object):
        FUSION_TYPE, EMBEDDING_NORMALIZATION, EMBEDDING_METRIC = "cat", False, "euclidean"

    class _Cfg(object):
        INPUT, TRAIN = "RGBD", _T

    saved = dict(N.CONFIG)
    try:
        N.configure(_Cfg)
        net = N.seg_resnet34_8s_embedding(2, 64, None)
        assert (net.fusion_type, net.normalize, net.feature_dim) == ("cat", False, 128)
        N.configure(_Cfg, live=True)                             # what shim.install() does: read at construction time
        _T.FUSION_TYPE = "add"
        assert N.seg_resnet34_8s_embedding(2, 64, None).fusion_type == "add"
        TD._LIVE_CFG[0] = _Cfg
        assert TD._metric() == "euclidean" and TD._metric("cosine") == "cosine"
    finally:
        N.CONFIG.update(saved)
        N._LIVE_CFG[0] = None
        TD._LIVE_CFG[0] = None
    assert TD._metric() == "cosine"
    # a checkpoint of another variant is filtered by name and shape like SEG.py:152 (nothing matches -> random init kept)
    sd = N.random_state_dict(64, seed=1, input_type="RGBD", fusion_type="early", in_channels=6)
    net = N.seg_resnet34_8s_embedding(2, 64, {"module." + k: v for k, v in sd.items()})
    assert not torch.equal(net.state_dict()["fcn.resnet34_8s.conv1.weight"][:, :3], sd["fcn.resnet34_8s.conv1.weight"][:, :3])
    assert torch.equal(net.state_dict()["fcn.resnet34_8s.layer1.0.conv1.weight"], sd["fcn.resnet34_8s.layer1.0.conv1.weight"])


def test_shim_rebinds_the_reference_call_sites_and_reads_its_cfg_live():
    """shim.install(): every reference call site of the path points at this package afterwards, and the factories /
    clustering_features read INPUT, FUSION_TYPE, EMBEDDING_NORMALIZATION, EMBEDDING_METRIC from the reference's own cfg
    object at construction / call time (the tools call cfg_from_file() after the imports)."""
    import sys
    import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present (GPU box)")
    from unseenobjectclustering_b200 import shim, mean_shift as MS, evaluation as EV
    ref = rh.load()
    mods = [sys.modules[n] for n in ("networks", "fcn.test_dataset", "utils.mean_shift", "utils.evaluation")]
    saved = [dict(m.__dict__) for m in mods]
    cfg_saved = (ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE, ref.cfg.TRAIN.EMBEDDING_METRIC)
    try:
        patched = shim.install()
        assert ref.networks.__dict__["seg_resnet34_8s_embedding"] is N.seg_resnet34_8s_embedding
        assert ref.networks.__dict__["seg_resnet34_8s_embedding_early"] is N.seg_resnet34_8s_embedding_early
        for name in ("clustering_features", "crop_rois", "match_label_crop", "filter_labels_depth", "test_sample"):
            assert getattr(ref.test_dataset, name) is getattr(TD, name), name
        assert ref.test_dataset.multilabel_metrics is EV.multilabel_metrics
        assert ref.mean_shift.mean_shift_smart_init is MS.mean_shift_smart_init
        assert any("cfg" in p for p in patched)
        ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE = "COLOR", "add"          # as cfg_from_file() would, after install()
        net = ref.networks.__dict__["seg_resnet34_8s_embedding"](2, 64, None)
        assert net.input_type == "COLOR" and len(net.state_dict()) == 218
        ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE = "RGBD", "cat"
        assert ref.networks.__dict__["seg_resnet34_8s_embedding"](2, 64, None).feature_dim == 128
        ref.cfg.TRAIN.EMBEDDING_METRIC = "euclidean"
        assert TD._metric() == "euclidean"
    finally:
        ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE, ref.cfg.TRAIN.EMBEDDING_METRIC = cfg_saved
        for m, d in zip(mods, saved):
            for k, v in d.items():
                if m.__dict__.get(k) is not v:
                    m.__dict__[k] = v
        N._LIVE_CFG[0] = None
        TD._LIVE_CFG[0] = None
        N.CONFIG.update({"INPUT": "RGBD", "FUSION_TYPE": "add", "EMBEDDING_NORMALIZATION": True})
    assert ref.test_dataset.clustering_features is not TD.clustering_features


def test_metrics_host_half_matches_oracle_from_exact_counts():
    """evaluation.metrics_from_counts (Hungarian matching + the reference's float64 ratios) on count tables computed
    here with the oracle's primitives -- the same tables the device kernels produce (tests/test_gpu_pipeline.py)."""
    from unseenobjectclustering_b200 import evaluation as EV
    for seed in range(3):
        _, gt = O.synthetic_clustered_features(60, 76, 8, 3 + seed, 0.05, 600 + seed)
        gt = gt.numpy().astype(np.int64)
        pred = O.synthetic_prediction(gt, seed)
        tp = np.zeros((256, 256), dtype=np.int64)
        bp = np.zeros((256, 256), dtype=np.int64)
        br = np.zeros((256, 256), dtype=np.int64)
        np.add.at(tp, (gt.ravel(), pred.ravel()), 1)
        for i in np.unique(gt):
            for j in np.unique(pred):
                if i and j:
                    bp[i, j], br[i, j] = O.boundary_overlap(pred == j, gt == i)
        dp = sum(int(O.seg2bmap(pred == j).sum()) for j in np.unique(pred) if j)
        dg = sum(int(O.seg2bmap(gt == i).sum()) for i in np.unique(gt) if i)
        got = EV.metrics_from_counts(tp, bp, br, (dp, dg))
        want = O.multilabel_metrics(pred.astype(np.float32), gt.astype(np.float32))
        assert set(got) == set(want)
        for k in want:
            assert abs(float(got[k]) - float(want[k])) < 1e-12, (seed, k)
    # degenerate cases (evaluation.py:139-174)
    z = np.zeros((256, 256), dtype=np.int64)
    only_gt = z.copy(); only_gt[3, 0] = 50; only_gt[0, 0] = 100
    assert EV.metrics_from_counts(only_gt, z, z, (0, 10))['Objects Recall'] == 0.
    only_pred = z.copy(); only_pred[0, 4] = 50; only_pred[0, 0] = 100
    assert EV.metrics_from_counts(only_pred, z, z, (10, 0))['Objects Precision'] == 0.
    none = z.copy(); none[0, 0] = 100
    assert EV.metrics_from_counts(none, z, z, (0, 0))['obj_detected_075_percentage'] == 1.
