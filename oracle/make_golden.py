"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_harness.py) on seeded synthetic inputs.  TEST INFRASTRUCTURE ONLY; run in the
build container:  python oracle/make_golden.py

The reference ships no golden vectors of its own, so these fixtures -- outputs of
the reference itself -- are what pins the oracle and the CUDA path.  Inputs are regenerated from
seeds by oracle/uoc_oracle.py's generators (deterministic torch CPU RNG), so the fixtures only store
the seeds, small inputs and the reference outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh          # noqa: E402
import uoc_oracle as O            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CLUSTER_CASES = [
    # name, H, W, d, objects, noise, seed, num_seeds, np_seed
    ("cluster_a", 32, 48, 64, 4, 0.05, 11, 100, 3),
    ("cluster_b", 40, 40, 64, 6, 0.04, 12, 100, 5),
    ("cluster_c", 30, 44, 128, 3, 0.05, 13, 100, 7),
    ("cluster_d", 24, 24, 64, 2, 0.05, 14, 40, 9),
]


def gen_cluster(ref):
    for name, H, W, d, K, noise, seed, m, npseed in CLUSTER_CASES:
        feats, gt = O.synthetic_clustered_features(H, W, d, K, noise, seed)
        X = feats[0].view(d, -1).t()
        np.random.seed(npseed)
        first = np.random.randint(0, H * W)
        np.random.seed(npseed)
        labels, selected = ref.mean_shift.mean_shift_smart_init(X, kappa=20, num_seeds=m, max_iters=10, metric='cosine')
        # intermediates through the reference's own sub-functions
        np.random.seed(npseed)
        seeds, sel2 = ref.mean_shift.select_smart_seeds(X, m, return_selected_indices=True, metric='cosine')
        assert torch.equal(sel2, selected)
        seed_labels, Z = ref.mean_shift.mean_shift_with_seeds(X, seeds, 20, max_iters=10, metric='cosine')
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, d=d, objects=K, noise=noise, seed=seed,
                            num_seeds=m, first_index=first, features=feats.numpy(), gt=gt.numpy(),
                            labels=labels.numpy(), selected=selected.numpy(), Z=Z.numpy(),
                            seed_labels=seed_labels.numpy())
        print(name, "labels", np.unique(labels.numpy()).tolist())


EUCLID_CASES = [
    # name, H, W, d, objects, noise, seed, num_seeds, np_seed, scale (features are multiplied by it: not unit norm)
    ("euclid_a", 32, 48, 64, 4, 0.02, 41, 100, 3, 1.0),
    ("euclid_b", 30, 40, 64, 3, 0.03, 42, 60, 5, 1.5),
    ("euclid_c", 24, 32, 128, 2, 0.02, 43, 100, 7, 1.0),
]


def gen_euclid(ref):
    """metric='euclidean' (cfg.TRAIN.EMBEDDING_METRIC default, lib/fcn/config.py:261) through the
    reference's own mean_shift functions."""
    for name, H, W, d, K, noise, seed, m, npseed, scale in EUCLID_CASES:
        feats, gt = O.synthetic_clustered_features(H, W, d, K, noise, seed)
        feats = feats * scale
        X = feats[0].view(d, -1).t()
        np.random.seed(npseed)
        first = np.random.randint(0, H * W)
        np.random.seed(npseed)
        labels, selected = ref.mean_shift.mean_shift_smart_init(X, kappa=20, num_seeds=m, max_iters=10, metric='euclidean')
        np.random.seed(npseed)
        seeds, sel2 = ref.mean_shift.select_smart_seeds(X, m, return_selected_indices=True, metric='euclidean')
        assert torch.equal(sel2, selected)
        seed_labels, Z = ref.mean_shift.mean_shift_with_seeds(X, seeds, 20, max_iters=10, metric='euclidean')
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, d=d, objects=K, noise=noise, seed=seed, scale=scale,
                            num_seeds=m, first_index=first, features=feats.numpy(), gt=gt.numpy(),
                            labels=labels.numpy(), selected=selected.numpy(), Z=Z.numpy(),
                            seed_labels=seed_labels.numpy())
        print(name, "labels", np.unique(labels.numpy()).tolist(), "seed labels", len(np.unique(seed_labels.numpy())))


METRIC_CASES = [(96, 128, 4, 0), (96, 128, 5, 1), (75, 101, 4, 2), (120, 160, 6, 3), (480, 640, 6, 4), (64, 64, 3, 5)]
METRIC_KEYS = ['Objects F-measure', 'Objects Precision', 'Objects Recall', 'Boundary F-measure', 'Boundary Precision',
               'Boundary Recall', 'obj_detected', 'obj_detected_075', 'obj_gt', 'obj_detected_075_percentage']


def gen_metrics(ref):
    """utils.evaluation.multilabel_metrics of the unmodified reference (lib/utils/evaluation.py:109-257; skimage's disk
    supplied by ref_harness) on synthetic (prediction, gt) pairs; plus its three degenerate cases."""
    out = {"cases": len(METRIC_CASES)}
    for k, (H, W, K, seed) in enumerate(METRIC_CASES):
        _, gt = O.synthetic_clustered_features(H, W, 8, K, 0.05, 200 + seed)
        gt = gt.numpy().astype(np.float32)
        pred = O.synthetic_prediction(gt.astype(np.int64), seed).astype(np.float32)
        m = ref.evaluation.multilabel_metrics(pred, gt)
        out["meta%d" % k] = np.array([H, W, K, seed])
        out["values%d" % k] = np.array([float(m[key]) for key in METRIC_KEYS])
        print("metrics", k, {key: round(float(m[key]), 4) for key in METRIC_KEYS[:6]})
    z = np.zeros((40, 48), dtype=np.float32)
    one = z.copy(); one[5:20, 6:30] = 3
    for name, (p, g) in (("none_pred", (z, one)), ("none_gt", (one, z)), ("none_both", (z, z))):
        m = ref.evaluation.multilabel_metrics(p, g)
        out[name] = np.array([float(m[key]) for key in METRIC_KEYS])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)


def gen_two_stage(ref):
    H, W = 96, 128
    feats, gt = O.synthetic_clustered_features(H, W, 64, 4, 0.05, seed=21)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=22)
    xyz[:, 2, :10, :] = 0          # some invalid depth so that filter_labels_depth has something to see

    def net(i, l, dd):
        return feats

    def net_crop(i, l, dd):
        return torch.cat([O.synthetic_clustered_features(224, 224, 64, 2, 0.05, seed=30 + k)[0] for k in range(i.shape[0])], 0)

    np.random.seed(3)
    n1 = [np.random.randint(0, H * W)]
    np.random.seed(3)
    with rh.cpu_cuda_identity():
        rgb_c, mask_c, rois, depth_c = None, None, None, None
        out_label, refined = ref.test_dataset.test_sample({'image_color': img, 'depth': xyz}, net, net_crop)
    # also record the crop stage in isolation (reference functions)
    rgb_c, mask_c, rois, depth_c = ref.test_dataset.crop_rois(img, out_label.clone(), xyz)
    np.random.seed(3)
    np.random.randint(0, H * W)
    firsts_crop = [int(np.random.randint(0, 224 * 224)) for _ in range(rgb_c.shape[0])]
    np.savez_compressed(os.path.join(OUT, "two_stage.npz"), H=H, W=W, feat_seed=21, frame_seed=22, crop_seed0=30,
                        first_index=n1[0], first_indices_crop=np.array(firsts_crop), out_label=out_label.numpy(),
                        refined=refined.numpy(), rois=rois.numpy(), mask_crops_sum=mask_c.sum((1, 2)).numpy(),
                        rgb_crops_mean=rgb_c.mean((1, 2, 3)).numpy(), depth_crops_mean=depth_c.mean((1, 2, 3)).numpy())
    print("two_stage", np.unique(out_label.numpy()).tolist(), np.unique(refined.numpy()).tolist(), rois.numpy().tolist())


def gen_backbone(ref):
    from unseenobjectclustering_b200.networks import random_state_dict
    for name, H, W, wseed, fseed in (("backbone_a", 64, 96, 5, 6), ("backbone_b", 48, 80, 7, 8)):
        sd = O.randomise_bn_(random_state_dict(64, seed=wseed), wseed + 1000)   # non-trivial BN: exercises the folding
        net = rh.build_network(64, state_dict=sd)
        missing = [k for k, v in net.state_dict().items() if k in sd and not torch.equal(v, sd[k])]
        assert not missing, missing[:5]
        img, xyz = O.synthetic_rgbd_frame(H, W, seed=fseed)
        with torch.no_grad():
            feats = net(img, None, xyz)
            ta = net.fcn.resnet34_8s(img)
            tb = net.fcn_depth.resnet34_8s(xyz)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, weight_seed=wseed, frame_seed=fseed,
                            trunk_rgb=ta.numpy(), trunk_depth=tb.numpy(), features_sub=feats[:, :, ::4, ::4].numpy(),
                            features_sum=feats.double().sum().item())
        print(name, ta.shape, float(ta.abs().mean()), float(tb.abs().mean()))


VARIANT_CASES = [
    # name, factory, input_type, fusion_type, normalize, in_channels, H, W, weight seed, frame seed
    ("variant_color", "seg_resnet34_8s_embedding", "COLOR", "add", True, 3, 48, 64, 21, 22),
    ("variant_depth", "seg_resnet34_8s_embedding", "DEPTH", "add", True, 3, 48, 64, 23, 24),
    ("variant_early", "seg_resnet34_8s_embedding_early", "RGBD", "early", True, 6, 48, 64, 25, 26),
    ("variant_cat", "seg_resnet34_8s_embedding", "RGBD", "cat", True, 3, 48, 64, 27, 28),
    ("variant_add_nonorm", "seg_resnet34_8s_embedding", "RGBD", "add", False, 3, 48, 64, 29, 30),
]


def gen_variants(ref):
    """The reference's other input / fusion variants (SEG.py:97-114), built through its own
    factories with cfg.INPUT / cfg.TRAIN.FUSION_TYPE / cfg.TRAIN.EMBEDDING_NORMALIZATION set."""
    from unseenobjectclustering_b200.networks import random_state_dict
    for name, factory, it, ft, norm, cin, H, W, wseed, fseed in VARIANT_CASES:
        sd = O.randomise_bn_(random_state_dict(64, seed=wseed, input_type=it, fusion_type=ft, in_channels=cin), wseed + 1000)
        net = rh.build_network(64, state_dict=sd, name=factory, input_type=it, fusion_type=ft, normalize=norm)
        have = net.state_dict()
        assert sorted(have.keys()) == sorted(sd.keys()), (name, set(have) ^ set(sd))
        assert all(torch.equal(have[k], sd[k]) for k in sd), name
        img, xyz = O.synthetic_rgbd_frame(H, W, seed=fseed)
        with torch.no_grad():
            feats = net(img, None, xyz)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, weight_seed=wseed, frame_seed=fseed,
                            input_type=it, fusion_type=ft, normalize=norm, in_channels=cin, factory=factory,
                            features_sub=feats[:, :, ::2, ::2].numpy(), features_sum=feats.double().sum().item())
        print(name, tuple(feats.shape), float(feats.abs().mean()))


def gen_input_prep(ref):
    """The reference's own read_sample (tools/test_images.py:105-135) on windows of its demo frames (data/demo):
    the windows are written as PNG files so that the unmodified function -- cv2.imread included -- produces the
    expected tensors.  compute_xyz uses absolute pixel indices, so a window is a small frame of its own."""
    import json
    import tempfile
    import cv2
    tool = rh.load_tool("test_images")
    demo = os.path.join(rh.REF_ROOT, "data", "demo")
    cam = json.load(open(os.path.join(demo, "camera_params.json")))
    cases = [("000000", 150, 200, 40, 56), ("000003", 0, 0, 33, 47), ("000006", 300, 420, 48, 64)]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for k, (stem, y0, x0, h, w) in enumerate(cases):
            im = cv2.imread(os.path.join(demo, stem + "-color.png"))[y0:y0 + h, x0:x0 + w].copy()
            dp = cv2.imread(os.path.join(demo, stem + "-depth.png"), cv2.IMREAD_ANYDEPTH)[y0:y0 + h, x0:x0 + w].copy()
            if k == 1:
                dp[5:9, 7:20] = 0                     # a hole: invalid depth stays exactly 0 in all three channels
            fc, fd = os.path.join(tmp, "c%d.png" % k), os.path.join(tmp, "d%d.png" % k)
            cv2.imwrite(fc, im)
            cv2.imwrite(fd, dp)
            sample = tool.read_sample(fc, fd, cam)
            out["im%d" % k] = im
            out["depth%d" % k] = dp
            out["image_color%d" % k] = sample["image_color"].numpy()
            out["xyz%d" % k] = sample["depth"].numpy()
            print("input_prep", k, im.shape, dp.dtype, float(sample["depth"].abs().mean()))
    np.savez_compressed(os.path.join(OUT, "input_prep.npz"), cases=len(cases), fx=cam["fx"], fy=cam["fy"],
                        x_offset=cam["x_offset"], y_offset=cam["y_offset"], **out)


FULL_CASES = [
    # name, H, W, d, objects, noise, first generator seed to try, num_seeds, max_iters, np_seed
    ("full_cfg2", 480, 640, 64, 6, 0.05, 100, 100, 10, 3),          # BASELINE config 2
    ("full_cfg5", 720, 960, 128, 12, 0.05, 500, 100, 30, 3),        # BASELINE config 5 (one GPU's share)
]


INIT_CASES = [
    # name, H, W, d, objects, noise, seed, num_seeds, num_init, metric, scale
    ("init_a", 32, 48, 64, 4, 0.05, 51, 60, 5, "cosine", 1.0),
    ("init_b", 30, 40, 64, 3, 0.03, 52, 40, 12, "euclidean", 1.5),
]


def gen_init(ref):
    """select_smart_seeds(init_seeds=, num_init_seeds=) of the unmodified reference (lib/utils/mean_shift.py:144-149, :164-169):
    the given seeds are the object centres' first rows plus noise (NOT points of X)."""
    for name, H, W, d, K, noise, seed, m, k, metric, scale in INIT_CASES:
        feats, gt = O.synthetic_clustered_features(H, W, d, K, noise, seed)
        feats = feats * scale
        X = feats[0].view(d, -1).t().contiguous()
        g = torch.Generator().manual_seed(seed + 1000)
        given = torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=1) * scale
        init = torch.zeros((m, d))
        init[:k] = given
        seeds, selected = ref.mean_shift.select_smart_seeds(X, m, return_selected_indices=True, init_seeds=init,
                                                            num_init_seeds=k, metric=metric)
        assert seeds is init
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, d=d, objects=K, noise=noise, seed=seed, scale=scale,
                            num_seeds=m, num_init=k, metric=metric, given=given.numpy(), selected=selected.numpy(),
                            seeds=seeds.numpy())
        print(name, selected[:k + 4].tolist())


def gen_munkres(ref):
    """Assignments of the reference's own utils.munkres.Munkres on tie-heavy rectangular cost matrices (the way
    lib/utils/evaluation.py:220-221 calls it: F.max() - F)."""
    import importlib
    mk = importlib.import_module("utils.munkres")
    rng = np.random.default_rng(1234)
    mats, shapes, pairs, counts = [], [], [], []
    for t in range(400):
        r, c = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        mode = t % 4
        if mode == 0:
            F = rng.random((r, c))
        elif mode == 1:
            F = rng.integers(0, 4, (r, c)).astype(np.float64) / 3.0
        elif mode == 2:
            F = np.round(rng.random((r, c)), 1)
            F[rng.random((r, c)) < 0.5] = 0
        else:
            F = rng.integers(0, 2, (r, c)).astype(np.float64)
        cost = F.max() - F.copy()
        a = mk.Munkres().compute(cost.copy())
        pad = np.zeros((9, 9))
        pad[:r, :c] = cost
        mats.append(pad)
        shapes.append((r, c))
        pa = -np.ones((9, 2), dtype=np.int64)
        pa[:len(a)] = np.array(a, dtype=np.int64).reshape(-1, 2)
        pairs.append(pa)
        counts.append(len(a))
    np.savez_compressed(os.path.join(OUT, "munkres.npz"), cost=np.stack(mats), shape=np.array(shapes), pairs=np.stack(pairs),
                        count=np.array(counts))
    print("munkres", len(mats), "matrices")


def _planar_to_X(Xp):
    """[d, n] planar float32 -> the reference's X: a [n, d] view with strides (1, n) (test_dataset.py:54-55)."""
    return torch.from_numpy(Xp).t()


def _margin_seed(ref, H, W, d, K, noise, seed0, m, first, tries=20):
    """First generator seed >= seed0 whose farthest-point sequence is the same in the reference (torch CPU, library
    summation order) and in the canonical-order C oracle: such an input has decision margins larger than fp32
    summation-order noise, so ANY correct fp32 implementation must reproduce all m indices (goldens are exact
    only for integer outputs on inputs with margin)."""
    import uoc_oracle_c as OC
    for seed in range(seed0, seed0 + tries):
        Xp, gt = O.exact_clustered_field(H, W, d, K, noise, seed)
        X = _planar_to_X(Xp)
        seeds, sel = _ref_select(ref, X, m, first)
        sel_c, _ = OC.select_seeds(Xp, m, first)
        if np.array_equal(sel.numpy(), sel_c):
            return seed, Xp, gt
        print("   seed", seed, "has a marginless farthest-point decision (reference != canonical order); next")
    raise RuntimeError("no seed with margin found")


def _ref_select(ref, X, m, first):
    """The reference's select_smart_seeds with its np.random.randint(0, n) draw (mean_shift.py:155) pinned to `first`."""
    saved = np.random.randint
    np.random.randint = lambda *a, **k: first
    try:
        return ref.mean_shift.select_smart_seeds(X, m, return_selected_indices=True, metric='cosine')
    finally:
        np.random.randint = saved


def _ref_cluster(ref, X, m, iters, first):
    saved = np.random.randint
    np.random.randint = lambda *a, **k: first
    try:
        labels, selected = ref.mean_shift.mean_shift_smart_init(X, kappa=20, num_seeds=m, max_iters=iters, metric='cosine')
        seeds, sel2 = ref.mean_shift.select_smart_seeds(X, m, return_selected_indices=True, metric='cosine')
        assert torch.equal(sel2, selected)
        seed_labels, Z = ref.mean_shift.mean_shift_with_seeds(X, seeds, 20, max_iters=iters, metric='cosine')
    finally:
        np.random.randint = saved
    return labels, selected, seed_labels, Z


def gen_full(ref):
    """BASELINE-size fixtures: the UNMODIFIED reference on 640x480x64 (config 2) and 960x720x128 with
    30 updates (config 5).  Only the outputs are stored (labels as uint8); the inputs are regenerated bit-identically from
    the seed by O.exact_clustered_field and checked by CRC."""
    for name, H, W, d, K, noise, seed0, m, iters, npseed in FULL_CASES:
        np.random.seed(npseed)
        first = int(np.random.randint(0, H * W))
        seed, Xp, gt = _margin_seed(ref, H, W, d, K, noise, seed0, m, first)
        X = _planar_to_X(Xp)
        labels, selected, seed_labels, Z = _ref_cluster(ref, X, m, iters, first)
        assert int(labels.max()) < 256
        np.savez_compressed(os.path.join(OUT, name + ".npz"), H=H, W=W, d=d, objects=K, noise=noise, seed=seed, num_seeds=m,
                            max_iters=iters, first_index=first, crc32=O.field_crc32(Xp), gt=gt.astype(np.uint8),
                            labels=labels.numpy().astype(np.uint8), selected=selected.numpy(), Z=Z.numpy(),
                            seed_labels=seed_labels.numpy())
        print(name, "seed", seed, "labels", np.unique(labels.numpy(), return_counts=True))


def gen_full_two_stage(ref):
    """BASELINE config 3 at full size: the reference's own test_sample (lib/fcn/test_dataset.py:232-267) on a 640x480
    frame whose stage-1 field has 6 objects, 224x224 crop fields from the same generator (fake networks, as SURVEY 8c
    advises: random-init embeddings collapse)."""
    H, W, K = 480, 640, 6
    np.random.seed(3)
    first = int(np.random.randint(0, H * W))
    seed, Xp, gt = _margin_seed(ref, H, W, 64, K, 0.05, 300, 100, first)
    feats = torch.from_numpy(Xp).view(1, 64, H, W)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=22)
    xyz[:, 2, :40, :] = 0          # invalid depth on top: filter_labels_depth has something to see

    # crops: one field per crop; first indices pinned; each checked for margin like the stage-1 field
    firsts_crop, crop_seeds, crop_fields = [], [], []
    rs = np.random.RandomState(5)
    for k in range(K):
        fc = int(rs.randint(0, 224 * 224))
        sc, Xc, _ = _margin_seed(ref, 224, 224, 64, 2, 0.05, 700 + 20 * k, 100, fc)
        firsts_crop.append(fc); crop_seeds.append(sc); crop_fields.append(torch.from_numpy(Xc).view(1, 64, 224, 224))

    def net(i, l, dd):
        return feats

    def net_crop(i, l, dd):
        return torch.cat(crop_fields[:i.shape[0]], 0)

    draws = [first] + firsts_crop
    saved = np.random.randint
    np.random.randint = lambda *a, **k: draws.pop(0)
    try:
        with rh.cpu_cuda_identity():
            out_label, refined = ref.test_dataset.test_sample({'image_color': img, 'depth': xyz}, net, net_crop)
    finally:
        np.random.randint = saved
    rgb_c, mask_c, rois, depth_c = ref.test_dataset.crop_rois(img, out_label.clone(), xyz)
    np.savez_compressed(os.path.join(OUT, "full_cfg3.npz"), H=H, W=W, objects=K, seed=seed, frame_seed=22, first_index=first,
                        crc32=O.field_crc32(Xp), crop_seeds=np.array(crop_seeds), first_indices_crop=np.array(firsts_crop),
                        num_crops=rgb_c.shape[0], out_label=out_label.numpy().astype(np.uint8),
                        refined=refined.numpy().astype(np.uint8), rois=rois.numpy())
    print("full_cfg3 seed", seed, "crops", rgb_c.shape[0], np.unique(out_label.numpy()).tolist(), np.unique(refined.numpy()).tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = rh.load()
    only = sys.argv[1:]
    if not only or "cluster" in only:
        gen_cluster(ref)
    if not only or "two_stage" in only:
        gen_two_stage(ref)
    if not only or "backbone" in only:
        gen_backbone(ref)
    if not only or "input_prep" in only:
        gen_input_prep(ref)
    if not only or "variants" in only:
        gen_variants(ref)
    if not only or "euclid" in only:
        gen_euclid(ref)
    if not only or "metrics" in only:
        gen_metrics(ref)
    if not only or "init" in only:
        gen_init(ref)
    if not only or "munkres" in only:
        gen_munkres(ref)
    if not only or "full" in only:
        gen_full(ref)
    if not only or "full_two_stage" in only:
        gen_full_two_stage(ref)
