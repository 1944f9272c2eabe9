"""Pin the oracles: torch oracle == golden fixtures (outputs of the unmodified reference), torch
oracle == reference run live (only where /root/reference exists), C oracle == torch oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import uoc_oracle as O
import uoc_oracle_c as C
import ref_harness as rh
from conftest import GOLDEN

CLUSTER_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "cluster_*.npz")))


def _load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("path", CLUSTER_FIXTURES, ids=[os.path.basename(p) for p in CLUSTER_FIXTURES])
def test_torch_oracle_matches_golden_clustering(path):
    g = _load(path)
    feats = torch.from_numpy(g["features"])
    d = int(g["d"])
    X = feats[0].view(d, -1).t()
    labels, sel, seeds, Z, sl = O.mean_shift_smart_init(X, num_seeds=int(g["num_seeds"]), first_index=int(g["first_index"]),
                                                        return_all=True)
    assert np.array_equal(sel.numpy(), g["selected"])
    assert np.array_equal(labels.numpy(), g["labels"])           # bit-exact: same torch ops as the reference
    assert np.array_equal(sl.numpy(), g["seed_labels"])
    assert np.allclose(Z.numpy(), g["Z"], atol=1e-6)
    # the synthetic generator is deterministic
    f2, gt2 = O.synthetic_clustered_features(int(g["H"]), int(g["W"]), d, int(g["objects"]), float(g["noise"]), int(g["seed"]))
    assert torch.equal(f2, feats)
    assert O.labels_equal_up_to_permutation(labels.numpy(), gt2.numpy().ravel()) or int(g["objects"]) + 1 != len(np.unique(g["labels"]))


@pytest.mark.parametrize("path", CLUSTER_FIXTURES, ids=[os.path.basename(p) for p in CLUSTER_FIXTURES])
def test_c_oracle_matches_golden_clustering(path):
    g = _load(path)
    d = int(g["d"])
    Xp = g["features"][0].reshape(d, -1)
    res = C.cluster(Xp, int(g["num_seeds"]), int(g["first_index"]))
    # farthest point sampling: identical indices on these inputs (margins >> fp32 rounding)
    assert np.array_equal(res["selected"], g["selected"])
    cosd = 1.0 - (res["Z"] * g["Z"]).sum(1)
    assert np.abs(cosd).max() < 1e-5
    assert np.array_equal(res["seed_labels"], g["seed_labels"])
    assert np.array_equal(res["labels"], g["labels"])


EUCLID_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "euclid_*.npz")))


@pytest.mark.parametrize("path", EUCLID_FIXTURES, ids=[os.path.basename(p) for p in EUCLID_FIXTURES])
def test_euclidean_oracles_match_reference_golden(path):
    """metric='euclidean' (lib/utils/mean_shift.py:21-24,58-60,101-105,159-160,207-209): torch oracle bit-identical to
    the fixtures written by the unmodified reference, C oracle (canonical fp32 chain + sqrtf) identical in every
    discrete decision; euclid_b is not unit norm."""
    g = _load(path)
    d, m = int(g["d"]), int(g["num_seeds"])
    feats = torch.from_numpy(g["features"])
    X = feats[0].view(d, -1).t()
    labels, sel, seeds, Z, sl = O.mean_shift_smart_init(X, num_seeds=m, first_index=int(g["first_index"]), return_all=True,
                                                        metric="euclidean")
    assert np.array_equal(sel.numpy(), g["selected"])
    assert np.array_equal(labels.numpy(), g["labels"])
    assert np.array_equal(sl.numpy(), g["seed_labels"])
    assert np.allclose(Z.numpy(), g["Z"], atol=1e-6)
    res = C.cluster(g["features"][0].reshape(d, -1), m, int(g["first_index"]), metric="euclidean")
    assert np.array_equal(res["selected"], g["selected"])
    assert np.abs(res["Z"] - g["Z"]).max() < 2e-5
    assert np.array_equal(res["seed_labels"], g["seed_labels"])
    assert np.array_equal(res["labels"], g["labels"])


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_euclidean_torch_oracle_is_bit_identical_to_the_live_reference():
    ref = rh.load()
    feats, _ = O.synthetic_clustered_features(28, 36, 64, 3, 0.02, seed=77)
    X = (feats[0] * 1.25).view(64, -1).t()
    np.random.seed(11)
    lr, sr = ref.mean_shift.mean_shift_smart_init(X, kappa=20, num_seeds=50, max_iters=10, metric='euclidean')
    np.random.seed(11)
    lo, so = O.mean_shift_smart_init(X, num_seeds=50, metric="euclidean")
    assert torch.equal(lr, lo) and torch.equal(sr, so)


INIT_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "init_*.npz")))


def _init_case(g):
    d, m, k = int(g["d"]), int(g["num_seeds"]), int(g["num_init"])
    feats, _ = O.synthetic_clustered_features(int(g["H"]), int(g["W"]), d, int(g["objects"]), float(g["noise"]), int(g["seed"]))
    feats = feats * float(g["scale"])
    return feats, d, m, k, str(g["metric"])


@pytest.mark.parametrize("path", INIT_FIXTURES, ids=[os.path.basename(p) for p in INIT_FIXTURES])
def test_select_seeds_with_init_seeds_matches_reference_golden(path):
    """select_smart_seeds(init_seeds=, num_init_seeds=) (lib/utils/mean_shift.py:144-149, :164-169): torch oracle bit-identical
    to the fixture written by the unmodified reference; C oracle (canonical chain) picks the same indices."""
    g = _load(path)
    feats, d, m, k, metric = _init_case(g)
    X = feats[0].view(d, -1).t().contiguous()
    init = torch.zeros((m, d))
    init[:k] = torch.from_numpy(g["given"])
    seeds, sel = O.select_seeds_init(X, m, init, k, metric)
    assert seeds is init
    assert np.array_equal(sel.numpy(), g["selected"]) and (sel[:k] == -1).all()
    assert np.array_equal(seeds.numpy(), g["seeds"])
    sel_c, seeds_c = C.select_seeds_init(feats[0].reshape(d, -1).numpy(), m, g["given"], metric)
    assert np.array_equal(sel_c, g["selected"])
    assert np.array_equal(seeds_c, g["seeds"])


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_select_seeds_with_init_seeds_is_bit_identical_to_the_live_reference():
    ref = rh.load()
    feats, _ = O.synthetic_clustered_features(28, 36, 64, 3, 0.04, seed=91)
    X = feats[0].view(64, -1).t().contiguous()
    given = torch.nn.functional.normalize(torch.randn(7, 64, generator=torch.Generator().manual_seed(5)), dim=1)
    a, b = torch.zeros((30, 64)), torch.zeros((30, 64))
    a[:7], b[:7] = given, given
    sr, ir = ref.mean_shift.select_smart_seeds(X, 30, return_selected_indices=True, init_seeds=a, num_init_seeds=7, metric='cosine')
    so, io = O.select_seeds_init(X, 30, b, 7, "cosine")
    assert torch.equal(ir, io) and torch.equal(sr, so)


def test_assignment_matches_the_reference_munkres_golden():
    """The Hungarian step of the evaluation tail (lib/utils/evaluation.py:220-221): the product's host implementation and
    the oracle's restatement both reproduce the assignments the reference's own Munkres class returned on 400 tie-heavy
    rectangular matrices (fixture written by oracle/make_golden.py) -- the pairs, not just the optimum."""
    from unseenobjectclustering_b200 import evaluation as EV
    g = _load(os.path.join(GOLDEN, "munkres.npz"))
    for k in range(g["cost"].shape[0]):
        r, c = int(g["shape"][k][0]), int(g["shape"][k][1])
        cost = g["cost"][k][:r, :c]
        want = [tuple(int(v) for v in p) for p in g["pairs"][k][:int(g["count"][k])]]
        assert EV._assignment(cost.copy()) == want, k
        assert O.munkres_assignment(cost.copy()) == want, k


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_assignment_matches_the_live_reference_munkres():
    import importlib
    from unseenobjectclustering_b200 import evaluation as EV
    rh.load()
    mk = importlib.import_module("utils.munkres")
    rng = np.random.default_rng(7)
    for t in range(600):
        r, c = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        F = rng.integers(0, 3, (r, c)).astype(np.float64) / 2.0 if t % 2 else np.round(rng.random((r, c)), 1)
        cost = F.max() - F.copy()
        want = [tuple(int(v) for v in p) for p in mk.Munkres().compute(cost.copy())]
        assert EV._assignment(cost.copy()) == want
        assert O.munkres_assignment(cost.copy()) == want


def test_two_stage_oracle_matches_golden():
    g = _load(os.path.join(GOLDEN, "two_stage.npz"))
    H, W = int(g["H"]), int(g["W"])
    feats, _ = O.synthetic_clustered_features(H, W, 64, 4, 0.05, seed=int(g["feat_seed"]))
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=int(g["frame_seed"]))
    xyz[:, 2, :10, :] = 0

    def net(i, l, dd):
        return feats

    def net_crop(i, l, dd):
        return torch.cat([O.synthetic_clustered_features(224, 224, 64, 2, 0.05, seed=int(g["crop_seed0"]) + k)[0]
                          for k in range(i.shape[0])], 0)

    a, b = O.test_sample(img, xyz, net, net_crop, [int(g["first_index"])], g["first_indices_crop"].tolist())
    assert np.array_equal(a.numpy(), g["out_label"])
    assert np.array_equal(b.numpy(), g["refined"])
    rgb_c, mask_c, rois, depth_c = O.crop_rois(img, a.clone(), xyz)
    assert np.array_equal(rois.numpy(), g["rois"])
    assert np.allclose(mask_c.sum((1, 2)).numpy(), g["mask_crops_sum"])
    assert np.allclose(rgb_c.mean((1, 2, 3)).numpy(), g["rgb_crops_mean"], atol=1e-6)
    assert np.allclose(depth_c.mean((1, 2, 3)).numpy(), g["depth_crops_mean"], atol=1e-6)


@pytest.mark.parametrize("name", ["backbone_a", "backbone_b"])
def test_backbone_oracle_matches_golden(name):
    from unseenobjectclustering_b200.networks import random_state_dict
    g = _load(os.path.join(GOLDEN, name + ".npz"))
    sd = O.randomise_bn_(random_state_dict(64, seed=int(g["weight_seed"])), int(g["weight_seed"]) + 1000)
    img, xyz = O.synthetic_rgbd_frame(int(g["H"]), int(g["W"]), seed=int(g["frame_seed"]))
    net = O.OracleSegNet(sd)
    with torch.no_grad():
        ta = O.resnet34_8s_trunk(img, net.sd, "fcn.resnet34_8s.")
        tb = O.resnet34_8s_trunk(xyz, net.sd, "fcn_depth.resnet34_8s.")
    assert np.allclose(ta.numpy(), g["trunk_rgb"], atol=2e-5, rtol=1e-4)
    assert np.allclose(tb.numpy(), g["trunk_depth"], atol=2e-5, rtol=1e-4)
    f = net(img, None, xyz)
    assert np.allclose(f[:, :, ::4, ::4].numpy(), g["features_sub"], atol=1e-5)
    assert abs(f.double().sum().item() - float(g["features_sum"])) < 1e-2


def test_label_seeds_quirks():
    """Order dependence, overwrite of already-labelled seeds and label gaps."""
    def unit(v):
        v = np.asarray(v, dtype=np.float32)
        return v / np.linalg.norm(v)
    # chain a - b - c where a~b and b~c within eps but a !~ c: greedy, not transitive closure
    ang = [0.0, 0.35, 0.70, 2.0]      # cosine distance 0.5(1-cos 0.35) = 0.0303 <= 0.04 ; 0.5(1-cos 0.7)=0.1176
    Z = np.stack([unit([np.cos(a), np.sin(a), 0, 0]) for a in ang])
    t = O.label_seeds(torch.from_numpy(Z), 0.04).numpy()
    c, uniq = C.label_seeds(Z, 0.04)
    assert np.array_equal(t, c)
    assert t.tolist() == [0, 0, 0, 1] or t.tolist() == [0, 0, 1, 2]
    assert uniq == len(np.unique(t))
    rng = np.random.RandomState(0)
    for trial in range(20):
        base = rng.randn(5, 16).astype(np.float32)
        Zr = base[rng.randint(0, 5, 60)] + 0.12 * rng.randn(60, 16).astype(np.float32)
        Zr /= np.linalg.norm(Zr, axis=1, keepdims=True)
        t = O.label_seeds(torch.from_numpy(Zr), 0.04).numpy()
        c, uniq = C.label_seeds(Zr, 0.04)
        # identical unless some pair sits within fp32 rounding of the threshold
        dots = Zr @ Zr.T
        margin = np.abs(0.5 * (1 - dots) - 0.04).min()
        if margin > 1e-6:
            assert np.array_equal(t, c), trial
            assert uniq == len(np.unique(t))


def test_assign_histogram_quirk_with_label_gaps():
    rng = np.random.RandomState(1)
    d, n, m = 16, 500, 6
    Z = rng.randn(m, d).astype(np.float32)
    Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    X = Z[rng.randint(0, m, n)] + 0.05 * rng.randn(n, d).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    seed_labels = np.array([0, 2, 2, 5, 5, 5], dtype=np.int32)       # gaps: unique = {0,2,5} -> num = 3
    t = O.assign_and_relabel(torch.from_numpy(X), torch.from_numpy(Z), torch.from_numpy(seed_labels.astype(np.int64))).numpy()
    c = C.assign(np.ascontiguousarray(X.T), Z, seed_labels, 3)
    assert np.array_equal(t, c)


VARIANTS = ["variant_color", "variant_depth", "variant_early", "variant_cat", "variant_add_nonorm"]


def _variant_setup(g):
    from unseenobjectclustering_b200.networks import random_state_dict
    it, ft, norm, cin = str(g["input_type"]), str(g["fusion_type"]), bool(g["normalize"]), int(g["in_channels"])
    sd = O.randomise_bn_(random_state_dict(64, seed=int(g["weight_seed"]), input_type=it, fusion_type=ft, in_channels=cin),
                         int(g["weight_seed"]) + 1000)
    img, xyz = O.synthetic_rgbd_frame(int(g["H"]), int(g["W"]), seed=int(g["frame_seed"]))
    return sd, img, xyz, it, ft, norm, cin


@pytest.mark.parametrize("name", VARIANTS)
def test_network_variant_oracle_matches_reference_golden(name):
    """COLOR / DEPTH / early / cat / un-normalised variants (SEG.py:97-114) against fixtures written by the
    unmodified reference factories (oracle/make_golden.py gen_variants)."""
    g = _load(os.path.join(GOLDEN, name + ".npz"))
    sd, img, xyz, it, ft, norm, _ = _variant_setup(g)
    f = O.OracleSegNet(sd, it, ft, norm)(img, None, xyz)
    assert f.shape[1] == (128 if ft == "cat" and it == "RGBD" else 64)
    assert np.allclose(f[:, :, ::2, ::2].numpy(), g["features_sub"], atol=1e-5, rtol=1e-5)
    assert abs(f.double().sum().item() - float(g["features_sum"])) < 1e-2 * max(1.0, abs(float(g["features_sum"])) * 1e-3)


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_variant_state_dict_keys_match_the_reference_factories():
    from unseenobjectclustering_b200.networks import reference_state_dict_keys
    for factory, it, ft, cin in (("seg_resnet34_8s_embedding", "COLOR", "add", 3),
                                 ("seg_resnet34_8s_embedding_early", "RGBD", "early", 6),
                                 ("seg_resnet34_8s_embedding", "RGBD", "cat", 3)):
        ref_sd = rh.build_network(64, name=factory, input_type=it, fusion_type=ft).state_dict()
        ours = reference_state_dict_keys(64, it, ft, cin)
        assert [k for k, _ in ours] == list(ref_sd.keys()), factory
        for k, shape in ours:
            assert tuple(ref_sd[k].shape) == tuple(shape), k


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_torch_oracle_is_bit_identical_to_the_live_reference():
    ref = rh.load()
    for seed, (H, W, K) in enumerate([(36, 52, 3), (48, 48, 5)]):
        feats, _ = O.synthetic_clustered_features(H, W, 64, K, 0.05, seed=100 + seed)
        X = feats[0].view(64, -1).t()
        np.random.seed(seed)
        lr, sr = ref.mean_shift.mean_shift_smart_init(X, kappa=20, num_seeds=100, max_iters=10, metric='cosine')
        np.random.seed(seed)
        lo, so = O.mean_shift_smart_init(X)
        assert torch.equal(lr, lo) and torch.equal(sr, so)
        np.random.seed(seed)
        a, _ = ref.test_dataset.clustering_features(feats, num_seeds=50)
        np.random.seed(seed)
        b, _ = O.clustering_features(feats, num_seeds=50)
        assert torch.equal(a, b)


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_state_dict_keys_match_the_reference_module():
    from unseenobjectclustering_b200.networks import reference_state_dict_keys
    net = rh.build_network(64)
    ref_sd = net.state_dict()
    ours = reference_state_dict_keys(64)
    assert [k for k, _ in ours] == list(ref_sd.keys())
    for k, shape in ours:
        assert tuple(ref_sd[k].shape) == tuple(shape), k


def test_input_prep_oracle_matches_reference_golden():
    """oracle read_sample_arrays / compute_xyz == the reference's own read_sample on windows of its demo frames (bit-exact)."""
    g = np.load(os.path.join(GOLDEN, "input_prep.npz"))
    cam = {"fx": float(g["fx"]), "fy": float(g["fy"]), "x_offset": float(g["x_offset"]), "y_offset": float(g["y_offset"])}
    for k in range(int(g["cases"])):
        image, xyz = O.read_sample_arrays(g["im%d" % k], g["depth%d" % k], cam)
        assert image.dtype == torch.float32 and xyz.dtype == torch.float32
        assert np.array_equal(image.numpy(), g["image_color%d" % k])
        assert np.array_equal(xyz.numpy(), g["xyz%d" % k])
    assert np.all(g["xyz1"][0, :, 5:9, 7:20] == 0)                     # invalid depth -> exact zeros in x, y and z


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_input_prep_oracle_matches_the_live_reference_tool():
    tool = rh.load_tool("test_images")
    rng = np.random.RandomState(0)
    depth = (rng.randint(0, 3000, (20, 30)).astype(np.float32)) / 1000.0
    want = tool.compute_xyz(depth, 612.937, 613.173, 322.549, 248.158, 20, 30)
    got = O.compute_xyz(depth, 612.937, 613.173, 322.549, 248.158, 20, 30)
    assert want.dtype == got.dtype == np.float32 and np.array_equal(want, got)



# ----------------------------------------------------------------------------------------------
# evaluation tail: utils.evaluation.multilabel_metrics
# ----------------------------------------------------------------------------------------------
def _metric_cases():
    import make_golden as MG
    g = _load(os.path.join(GOLDEN, "metrics.npz"))
    for k in range(int(g["cases"])):
        H, W, K, seed = (int(v) for v in g["meta%d" % k])
        _, gt = O.synthetic_clustered_features(H, W, 8, K, 0.05, 200 + seed)
        gt = gt.numpy().astype(np.float32)
        pred = O.synthetic_prediction(gt.astype(np.int64), seed).astype(np.float32)
        yield pred, gt, dict(zip(MG.METRIC_KEYS, g["values%d" % k]))
    z = np.zeros((40, 48), dtype=np.float32)
    one = z.copy(); one[5:20, 6:30] = 3
    for name, (p, q) in (("none_pred", (z, one)), ("none_gt", (one, z)), ("none_both", (z, z))):
        yield p, q, dict(zip(MG.METRIC_KEYS, g[name]))


def test_metrics_oracle_matches_reference_golden():
    for pred, gt, want in _metric_cases():
        if pred.shape[0] >= 480:
            continue                                   # the per-pair CPU loop of the full frame runs in the GPU test only
        got = O.multilabel_metrics(pred, gt)
        for k, v in want.items():
            assert abs(float(got[k]) - float(v)) < 1e-12, (k, got[k], v)


@pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")
def test_metrics_oracle_equals_the_live_reference():
    ref = rh.load()
    _, gt = O.synthetic_clustered_features(80, 100, 8, 5, 0.05, 321)
    gt = gt.numpy().astype(np.float32)
    pred = O.synthetic_prediction(gt.astype(np.int64), 7).astype(np.float32)
    a, b = ref.evaluation.multilabel_metrics(pred, gt), O.multilabel_metrics(pred, gt)
    assert set(a) == set(b) and all(abs(float(a[k]) - float(b[k])) < 1e-12 for k in a)


# ---------------------------------------------------------------------------------------------
# BASELINE-size fixtures (tests/golden/full_*.npz: outputs of the unmodified reference; inputs regenerated from the seed)
# ---------------------------------------------------------------------------------------------
def _full_field(g):
    Xp, gt = O.exact_clustered_field(int(g["H"]), int(g["W"]), int(g["d"]) if "d" in g else 64, int(g["objects"]),
                                     float(g["noise"]) if "noise" in g else 0.05, int(g["seed"]))
    assert O.field_crc32(Xp) == int(g["crc32"]), "exact_clustered_field is not bit-reproducible on this platform"
    return Xp, gt


def test_exact_generator_is_pinned():
    """A small field with a CRC committed here: the generator of the full-size fixtures gives the same bits everywhere."""
    Xp, gt = O.exact_clustered_field(24, 40, 64, 3, 0.05, 7)
    assert Xp.shape == (64, 960) and Xp.dtype == np.float32
    assert np.abs((Xp.astype(np.float64) ** 2).sum(0) - 1.0).max() < 2e-7
    assert O.field_crc32(Xp) == 0x15131BCC, hex(O.field_crc32(Xp))


def test_oracles_match_reference_golden_at_baseline_size_config2():
    """640x480x64, 100 seeds, 10 updates (BASELINE config 2): torch oracle bit-identical to the reference's output, C
    oracle identical in every discrete decision (all 100 indices, seed labels, 307 200 pixel labels)."""
    g = _load(os.path.join(GOLDEN, "full_cfg2.npz"))
    Xp, gt = _full_field(g)
    m, first = int(g["num_seeds"]), int(g["first_index"])
    X = torch.from_numpy(Xp).t()
    labels, sel, seeds, Z, sl = O.mean_shift_smart_init(X, num_seeds=m, max_iters=int(g["max_iters"]), first_index=first,
                                                        return_all=True)
    assert np.array_equal(sel.numpy(), g["selected"])
    assert np.array_equal(labels.numpy(), g["labels"].astype(np.int64))
    assert np.array_equal(sl.numpy(), g["seed_labels"])
    assert np.allclose(Z.numpy(), g["Z"], atol=1e-6)
    assert O.labels_equal_up_to_permutation(labels.numpy(), gt.ravel())
    res = C.cluster(Xp, m, first, iters=int(g["max_iters"]))
    assert np.array_equal(res["selected"], g["selected"])
    assert np.abs(1.0 - (res["Z"] * g["Z"]).sum(1)).max() < 1e-5
    assert np.array_equal(res["seed_labels"], g["seed_labels"])
    assert np.array_equal(res["labels"], g["labels"].astype(np.int32))


def test_c_oracle_matches_reference_golden_at_baseline_size_config5():
    """960x720x128, 30 updates (BASELINE config 5, one GPU's share): the discrete stages of the canonical-order C oracle
    against the reference's output -- all 100 farthest-point indices, and seed labels / 691 200 pixel labels from the
    reference's converged seeds (the 30-update double-precision loop of the C oracle is exercised at this size by the GPU
    test; make_golden.py pins the torch oracle at this size where the reference is importable)."""
    g = _load(os.path.join(GOLDEN, "full_cfg5.npz"))
    Xp, gt = _full_field(g)
    sel, seeds = C.select_seeds(Xp, int(g["num_seeds"]), int(g["first_index"]))
    assert np.array_equal(sel, g["selected"])
    sl, uniq = C.label_seeds(g["Z"], 0.04)
    assert np.array_equal(sl, g["seed_labels"])
    labels = C.assign(Xp, g["Z"], sl, uniq)
    assert np.array_equal(labels, g["labels"].astype(np.int32))
    assert O.labels_equal_up_to_permutation(labels, gt.ravel())
