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


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_filter_labels_depth_matches_oracle(seed):
    H, W = 60, 80
    lab = _labels(H, W, 5, seed).float()[None]
    _, xyz = O.synthetic_rgbd_frame(H, W, seed)
    g = torch.Generator().manual_seed(seed)
    xyz[:, 2][torch.rand(1, H, W, generator=g) < 0.3] = 0
    xyz[:, 2, : H // 2, : W // 2] = 0
    for thr in (0.5, 0.8):
        a = O.filter_labels_depth(lab, xyz, thr)
        b = TD.filter_labels_depth(lab, xyz, thr)
        assert torch.equal(a, b)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_crop_rois_matches_oracle(seed):
    H, W = 96, 128
    lab = _labels(H, W, 4, seed).float()[None]
    img, xyz = O.synthetic_rgbd_frame(H, W, seed)
    ra, ma, roa, da = O.crop_rois(img, lab.clone(), xyz)
    rb, mb, rob, db = TD.crop_rois(img, lab.clone(), xyz)
    assert torch.equal(roa, rob)
    assert torch.equal(ma, mb)
    assert torch.allclose(ra, rb, atol=1e-6) and torch.allclose(da, db, atol=1e-6)
    # no depth
    ra, ma, roa, da = O.crop_rois(img, lab.clone(), None)
    rb, mb, rob, db = TD.crop_rois(img, lab.clone(), None)
    assert da is None and db is None and torch.equal(roa, rob)


def test_crop_rois_empty():
    img, xyz = O.synthetic_rgbd_frame(32, 48, 0)
    r, m, rois, d = TD.crop_rois(img, torch.zeros(1, 32, 48), xyz)
    assert r.shape[0] == 0 and rois.shape == (0, 4)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_match_label_crop_matches_oracle(seed):
    H, W = 96, 128
    lab = _labels(H, W, 4, seed).float()[None]
    img, xyz = O.synthetic_rgbd_frame(H, W, seed)
    rgb_c, mask_c, rois, depth_c = O.crop_rois(img, lab.clone(), xyz)
    K = rgb_c.shape[0]
    crops = torch.stack([_labels(224, 224, 3, 50 + seed * 10 + k).float() for k in range(K)])
    a, la = O.match_label_crop(lab, crops.clone(), mask_c, rois, depth_c)
    b, lb = TD.match_label_crop(lab, crops.clone(), mask_c, rois, depth_c)
    assert torch.equal(a, b)
    assert torch.equal(la, lb)
    a, _ = O.match_label_crop(lab, crops.clone(), mask_c, rois, None)
    b, _ = TD.match_label_crop(lab, crops.clone(), mask_c, rois, None)
    assert torch.equal(a, b)


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
