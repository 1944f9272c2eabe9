"""GPU parity of the frame-level API (test_sample: stage-1 + depth filter + crop refine) against
the golden fixture written by the unmodified reference's test_sample."""
import os

import numpy as np
import pytest
import torch

import uoc_oracle as O
from conftest import GOLDEN
from unseenobjectclustering_b200 import _lib
from unseenobjectclustering_b200 import test_dataset as TD

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_two_stage_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "two_stage.npz"))
    H, W = int(g["H"]), int(g["W"])
    feats, _ = O.synthetic_clustered_features(H, W, 64, 4, 0.05, seed=int(g["feat_seed"]))
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=int(g["frame_seed"]))
    xyz[:, 2, :10, :] = 0
    fd = feats.to(DEV)

    def net(i, l, dd):
        return fd

    def net_crop(i, l, dd):
        return torch.cat([O.synthetic_clustered_features(224, 224, 64, 2, 0.05, seed=int(g["crop_seed0"]) + k)[0]
                          for k in range(i.shape[0])], 0).to(DEV)

    out_label, refined = TD.test_sample({'image_color': img, 'depth': xyz}, net, net_crop, [int(g["first_index"])],
                                        g["first_indices_crop"].tolist(), flags=_lib.FLAG_SYNC_CHECK)
    assert out_label.dtype == torch.float32 and out_label.device.type == "cpu" and out_label.shape == (1, H, W)
    assert refined is not None and refined.dtype == torch.float32 and refined.device.type == "cpu"
    assert O.labels_equal_up_to_permutation(out_label.numpy(), g["out_label"])
    assert O.labels_equal_up_to_permutation(refined.numpy(), g["refined"])
    assert np.array_equal(out_label.numpy(), g["out_label"])     # ids agree too on this input
    assert np.array_equal(refined.numpy(), g["refined"])


def test_no_objects_returns_none_refined():
    H, W = 64, 96
    feats, _ = O.synthetic_clustered_features(H, W, 64, 0, 0.02, seed=1)     # background only
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=2)
    fd = feats.to(DEV)
    out_label, refined = TD.test_sample({'image_color': img, 'depth': xyz}, lambda i, l, d: fd, lambda i, l, d: None,
                                        [5], None)
    assert float(out_label.abs().max()) == 0.0
    assert refined is None


@pytest.mark.parametrize("H,W", [(240, 320), (123, 170), (77, 101)])
def test_end_to_end_with_real_backbone_runs(H, W):
    """Random-init weights collapse to one cluster, so this only checks the plumbing
    end to end through both CUDA stages -- also at frame sizes that are not multiples of 8 / 4 / 2."""
    from unseenobjectclustering_b200 import networks as NW
    net = NW.seg_resnet34_8s_embedding(2, 64, None).cuda(0)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=0)
    out_label, refined = TD.test_sample({'image_color': img, 'depth': xyz}, net, None, [17], flags=_lib.FLAG_SYNC_CHECK)
    assert out_label.shape == (1, H, W) and refined is None


def test_frame_pipeline_equals_serial_calls():
    """Frames in flight on two streams (pipeline.py) give exactly the labels of one-at-a-time calls."""
    from unseenobjectclustering_b200 import mean_shift as MS
    from unseenobjectclustering_b200.pipeline import FramePipeline
    H, W = 96, 128
    fields = [O.synthetic_clustered_features(H, W, 64, 3, 0.05, seed=70 + k)[0].to(DEV) for k in range(5)]
    firsts = [5, 50, 500, 5000, 11]
    it = iter(fields)

    def net(i, l, d):
        return next(it).clone()

    pipe = FramePipeline(net, H, W, depth=2)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=0)
    outs = []
    for k in range(5):
        pipe.submit(img, xyz, firsts[k])
        if len(pipe.pending) == 2:
            outs.append(pipe.collect_one()[0].clone())
    outs.extend(o[0].clone() for o in pipe.drain())
    assert len(outs) == 5
    for k in range(5):
        want, _ = MS.cluster_fields(fields[k], 100, first_indices=[firsts[k]], flags=_lib.FLAG_SYNC_CHECK)
        assert torch.equal(outs[k].view(-1).to(torch.int32), want[0].cpu())


class _FieldNet(torch.nn.Module):
    """Stand-in network whose output depends on the input: image channel 0 carries a per-pixel cluster id."""

    def __init__(self, d=64, k=8, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("centres", torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=1))
        self.register_buffer("noise", 0.05 * torch.randn(d, 96 * 128, generator=g))

    def forward(self, img, label=None, depth=None):
        ids = img[0, 0].reshape(-1).long()
        f = self.centres[ids].t() + self.noise + 0.0 * depth[0, 0].reshape(1, -1)
        return torch.nn.functional.normalize(f, dim=0).reshape(1, -1, img.shape[2], img.shape[3]).contiguous()


def test_frame_pipeline_with_cuda_graphs_equals_serial_calls():
    """After the warm-up every slot replays two captured CUDA graphs per frame; results must not change."""
    from unseenobjectclustering_b200 import mean_shift as MS
    from unseenobjectclustering_b200.pipeline import FramePipeline
    H, W = 96, 128
    net = _FieldNet().to(DEV)
    frames = []
    for k in range(8):
        _, gt = O.synthetic_clustered_features(H, W, 8, 3 + k % 3, 0.05, seed=90 + k)
        img = torch.zeros(1, 3, H, W)
        img[0, 0] = gt.float()
        frames.append(img)
    xyz = torch.ones(1, 3, H, W)
    firsts = [3, 33, 333, 3333, 7, 77, 777, 7777]
    pipe = FramePipeline(net, H, W, depth=2)
    outs = []
    for k in range(8):
        pipe.submit(frames[k], xyz, firsts[k])
        if len(pipe.pending) == 2:
            outs.append(pipe.collect_one()[0].clone())
    outs.extend(o[0].clone() for o in pipe.drain())
    assert pipe.graph_error is None, pipe.graph_error
    assert all(s.graph_a is not None for s in pipe.slots)          # frames 3.. ran from graphs
    for k in range(8):
        feats = net(frames[k].to(DEV), None, xyz.to(DEV))
        want, _ = MS.cluster_fields(feats, 100, first_indices=[firsts[k]], flags=_lib.FLAG_SYNC_CHECK)
        assert torch.equal(outs[k].view(-1).to(torch.int32), want[0].cpu()), k
        assert len(torch.unique(want)) >= 3


@pytest.mark.parametrize("H,W", [(96, 128), (77, 101)])
def test_frame_pipeline_raw_frames_equal_prepared_frames(H, W):
    """submit_raw (uint8 BGR + uint16 depth, inputs built on the device) == submit with the reference's fp32 inputs
    (also at a frame size that is a multiple of nothing)."""
    from unseenobjectclustering_b200.pipeline import FramePipeline
    from unseenobjectclustering_b200 import networks
    net = networks.seg_resnet34_8s_embedding(2, 64, networks.random_state_dict(64, seed=2)).to(DEV)
    cam = {"fx": 120.0, "fy": 121.5, "x_offset": 63.2, "y_offset": 47.9}
    rng = np.random.RandomState(4)
    raws = [(rng.randint(0, 256, (H, W, 3)).astype(np.uint8), rng.randint(300, 1500, (H, W)).astype(np.uint16)) for _ in range(4)]
    firsts = [9, 99, 999, 1234]
    pipe_a = FramePipeline(net, H, W, depth=2)
    pipe_b = FramePipeline(net, H, W, depth=2)
    a, b = [], []
    for k, (im, dp) in enumerate(raws):
        io, xo = O.read_sample_arrays(im, dp, cam)                   # the reference's CPU tensors
        pipe_a.submit(io, xo, firsts[k])
        pipe_b.submit_raw(torch.from_numpy(im).pin_memory(), torch.from_numpy(dp.view(np.int16)).pin_memory(), cam, firsts[k])
        if len(pipe_a.pending) == 2:
            a.append(pipe_a.collect_one()[0].clone())
            b.append(pipe_b.collect_one()[0].clone())
    a.extend(o[0].clone() for o in pipe_a.drain())
    b.extend(o[0].clone() for o in pipe_b.drain())
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # a random-init network maps everything to one cluster, so also compare the inputs the slots were fed bit for bit
    for k in (2, 3):
        io, xo = O.read_sample_arrays(raws[k][0], raws[k][1], cam)
        slot = pipe_b.slots[k % 2]
        assert torch.equal(slot.img_dev.cpu(), io) and torch.equal(slot.xyz_dev.cpu(), xo)


class _BatchFieldNet(torch.nn.Module):
    """Batch-capable stand-in network: image channel 0 carries a per-pixel cluster id."""

    def __init__(self, d=64, k=8, hw=96 * 128, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("centres", torch.nn.functional.normalize(torch.randn(k, d, generator=g), dim=1))
        self.register_buffer("noise", 0.05 * torch.randn(d, hw, generator=g))

    def forward(self, img, label=None, depth=None):
        N, _, H, W = img.shape
        ids = img[:, 0].reshape(N, -1).long()
        f = self.centres[ids].permute(0, 2, 1) + self.noise.unsqueeze(0) + 0.0 * depth[:, 0].reshape(N, 1, -1)
        return torch.nn.functional.normalize(f, dim=1).reshape(N, -1, H, W).contiguous()


@pytest.mark.parametrize("B,depth", [(2, 2), (3, 1)])
def test_frame_pipeline_frames_per_slot_equals_serial_calls(B, depth):
    """frames_per_slot = B: B frames go through every kernel together (eager warm-up, then CUDA graphs); an
    incomplete last batch is flushed by drain().  Every frame's labels equal the one-at-a-time call's."""
    from unseenobjectclustering_b200 import mean_shift as MS
    from unseenobjectclustering_b200.pipeline import FramePipeline
    H, W = 96, 128
    net = _BatchFieldNet().to(DEV)
    nframes = 4 * B * depth + 1                          # the last slot launch is incomplete
    frames = []
    for k in range(nframes):
        _, gt = O.synthetic_clustered_features(H, W, 8, 3 + k % 3, 0.05, seed=190 + k)
        img = torch.zeros(1, 3, H, W)
        img[0, 0] = gt.float()
        frames.append(img)
    xyz = torch.ones(1, 3, H, W)
    firsts = [(37 * k + 5) % (H * W) for k in range(nframes)]
    pipe = FramePipeline(net, H, W, depth=depth, frames_per_slot=B)
    outs = []
    for k in range(nframes):
        pipe.submit(frames[k], xyz, firsts[k])
        while len(pipe.pending) >= depth:
            outs.extend(o.clone() for o in pipe.collect_one()[0])
    for res in pipe.drain():
        outs.extend(o.clone() for o in res[0])
    assert pipe.graph_error is None, pipe.graph_error
    assert all(s.graph_a is not None for s in pipe.slots)
    assert len(outs) >= nframes
    for k in range(nframes):
        feats = net(frames[k].to(DEV), None, xyz.to(DEV))
        want, _ = MS.cluster_fields(feats, 100, first_indices=[firsts[k]], flags=_lib.FLAG_SYNC_CHECK)
        assert torch.equal(outs[k].view(-1).to(torch.int32), want[0].cpu()), k


def test_batch_of_full_frames_equals_single_calls():
    """Two 640x480 fields in one library call: they do not fit the resident-slice sampler together, so the fields take
    turns there (same kernel, same indices); loop and label kernels run at batch 2.  Everything equals the single calls."""
    from unseenobjectclustering_b200 import mean_shift as MS
    fa, _ = O.synthetic_clustered_features(480, 640, 64, 6, 0.05, seed=301)
    fb, _ = O.synthetic_clustered_features(480, 640, 64, 4, 0.05, seed=302)
    both = torch.cat([fa, fb], 0).to(DEV)
    MS.register_bf16_copy(both, MS.pack_bf16(both))
    lab, sel, Z, sl = MS.cluster_fields(both, 100, first_indices=[1234, 99999], flags=_lib.FLAG_SYNC_CHECK, return_seeds=True)
    for j, (f, first) in enumerate(((fa, 1234), (fb, 99999))):
        one = f.to(DEV)
        MS.register_bf16_copy(one, MS.pack_bf16(one))
        l1, s1, Z1, sl1 = MS.cluster_fields(one, 100, first_indices=[first], flags=_lib.FLAG_SYNC_CHECK, return_seeds=True)
        assert torch.equal(sel[j], s1[0])
        assert torch.equal(sl[j], sl1[0])
        assert torch.equal(lab[j], l1[0])
        assert (1 - (Z[j] * Z1[0]).sum(-1)).abs().max().item() < 1e-6


# ----------------------------------------------------------------------------------------------
# two-stage plumbing kernels (csrc/refine.cu) against the oracle (= the reference's torch code)
# ----------------------------------------------------------------------------------------------
def _gt_labels(H, W, K, seed):
    _, gt = O.synthetic_clustered_features(H, W, 8, K, 0.05, seed)
    return gt


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_filter_labels_depth_kernel_matches_oracle(seed):
    H, W = 60, 80
    lab = torch.stack([_gt_labels(H, W, 5, seed).float(), _gt_labels(H, W, 3, seed + 10).float()])
    xyz = torch.cat([O.synthetic_rgbd_frame(H, W, seed)[1], O.synthetic_rgbd_frame(H, W, seed + 1)[1]])
    g = torch.Generator().manual_seed(seed)
    xyz[:, 2][torch.rand(2, H, W, generator=g) < 0.3] = 0
    xyz[0, 2, : H // 2, : W // 2] = 0
    for thr in (0.5, 0.8):
        want = O.filter_labels_depth(lab, xyz, thr)
        got = TD.filter_labels_depth(lab.to(DEV), xyz.to(DEV), thr)
        assert got.is_cuda and got.dtype == lab.dtype and torch.equal(got.cpu(), want)


@pytest.mark.parametrize("seed,H,W", [(0, 96, 128), (1, 96, 128), (2, 75, 101), (3, 480, 640)])
def test_crop_rois_kernel_matches_oracle(seed, H, W):
    lab = _gt_labels(H, W, 5, seed).float()[None]
    img, xyz = O.synthetic_rgbd_frame(H, W, seed)
    ra, ma, roa, da = O.crop_rois(img, lab.clone(), xyz)
    rb, mb, rob, db = TD.crop_rois(img.to(DEV), lab.to(DEV), xyz.to(DEV))
    assert rb.is_cuda and torch.equal(roa, rob.cpu())                     # ids order, boxes, half-to-even padding, clamping
    assert torch.equal(ma, mb.cpu())                                      # legacy nearest rule
    assert torch.allclose(ra, rb.cpu(), atol=2e-6, rtol=1e-6) and torch.allclose(da, db.cpu(), atol=2e-6, rtol=1e-6)
    rb, mb, rob, db = TD.crop_rois(img.to(DEV), lab.to(DEV), None)
    assert db is None and torch.equal(roa, rob.cpu())
    r0, m0, rois0, d0 = TD.crop_rois(img.to(DEV), torch.zeros(1, H, W, device=DEV), xyz.to(DEV))
    assert r0.shape[0] == 0 and rois0.shape == (0, 4)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("with_depth", [True, False])
def test_match_label_crop_kernel_matches_oracle(seed, with_depth):
    H, W = 96, 128
    lab = _gt_labels(H, W, 4, seed).float()[None]
    img, xyz = O.synthetic_rgbd_frame(H, W, seed)
    if seed == 1:
        xyz[:, 2, :, : W // 3] = 0                                        # crops with holes in the depth
    rgb_c, mask_c, rois, depth_c = O.crop_rois(img, lab.clone(), xyz)
    K = rgb_c.shape[0]
    # crop label maps: some clusters agree with the mask crop (kept), some do not (dropped)
    crops = torch.stack([torch.where(mask_c[k] > 0, _gt_labels(224, 224, 2, 50 + seed * 10 + k).float() + 1,
                                     (_gt_labels(224, 224, 3, 80 + seed * 10 + k).float() + 4) * float(k % 2)) for k in range(K)])
    want_ref, want_lc = O.match_label_crop(lab.clone(), crops.clone(), mask_c, rois, depth_c if with_depth else None)
    got_ref, got_lc = TD.match_label_crop(lab.clone(), crops.to(DEV), mask_c.to(DEV), rois.to(DEV),
                                          depth_c.to(DEV) if with_depth else None)
    assert not got_ref.is_cuda and torch.equal(got_ref, want_ref)
    assert torch.equal(got_lc.cpu(), want_lc)
    assert len(torch.unique(want_ref)) > 2



# ----------------------------------------------------------------------------------------------
# evaluation tail (csrc/metrics.cu + evaluation.py) against the reference's own numbers
# ----------------------------------------------------------------------------------------------
def test_multilabel_metrics_match_reference_golden():
    """utils.evaluation.multilabel_metrics: every entry of the dictionary equals the unmodified reference's
    (tests/golden/metrics.npz) to 1e-12 -- the counts are exact integers, the ratios float64 like the reference's."""
    import test_oracle as TO
    from unseenobjectclustering_b200 import evaluation as EV
    n = 0
    for pred, gt, want in TO._metric_cases():
        got = EV.multilabel_metrics(pred, gt)
        assert set(got) == set(want)
        for k, v in want.items():
            assert abs(float(got[k]) - float(v)) < 1e-12, (pred.shape, k, got[k], v)
        n += 1
    assert n == 9


def test_multilabel_counts_match_oracle_pairwise():
    """The device counts for EVERY (gt, prediction) label pair against the oracle's per-pair masks, boundary maps and
    cv2 dilations (not only the matched pairs the metrics use); tensors on the device as input."""
    from unseenobjectclustering_b200 import evaluation as EV
    _, gt = O.synthetic_clustered_features(90, 121, 8, 5, 0.05, 400)
    gt = gt.numpy().astype(np.int64)
    pred = O.synthetic_prediction(gt, 3)
    tp, bp, br, (dp, dg) = EV.multilabel_counts(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV))
    for i in np.unique(gt):
        for j in np.unique(pred):
            assert tp[i, j] == np.count_nonzero((gt == i) & (pred == j))
            if i and j:
                a, b = O.boundary_overlap(pred == j, gt == i)
                assert (bp[i, j], br[i, j]) == (a, b), (i, j)
    assert dp == sum(int(O.seg2bmap(pred == j).sum()) for j in np.unique(pred) if j)
    assert dg == sum(int(O.seg2bmap(gt == i).sum()) for i in np.unique(gt) if i)
    with pytest.raises(_lib.UocError):
        EV.multilabel_counts(torch.full((8, 8), 300).to(DEV), torch.zeros(8, 8).to(DEV))


def test_two_stage_plumbing_edge_cases():
    """Objects touching the frame border (clamped boxes), a single-pixel object (extent 0: no padding, a 1x1 crop),
    one-object frames, label maps without objects -- device kernels against the oracle."""
    H, W = 40, 56
    lab = torch.zeros(1, H, W)
    lab[0, 0:9, 0:13] = 2            # top-left corner: padding is clamped at 0
    lab[0, 30:40, 44:56] = 5         # bottom-right corner: clamped at H-1 / W-1
    lab[0, 20, 30] = 7               # a single pixel
    img, xyz = O.synthetic_rgbd_frame(H, W, 3)
    ra, ma, roa, da = O.crop_rois(img, lab.clone(), xyz)
    rb, mb, rob, db = TD.crop_rois(img.to(DEV), lab.to(DEV), xyz.to(DEV))
    assert roa.shape[0] == 3 and torch.equal(roa, rob.cpu()) and torch.equal(ma, mb.cpu())
    assert torch.allclose(ra, rb.cpu(), atol=2e-6, rtol=1e-6) and torch.allclose(da, db.cpu(), atol=2e-6, rtol=1e-6)
    crops = ma + 1                                          # cluster 2 = the object, cluster 1 = the rest (dropped)
    want_ref, want_lc = O.match_label_crop(lab.clone(), crops.clone(), ma, roa, da)
    got_ref, got_lc = TD.match_label_crop(lab.clone(), crops.to(DEV), mb, rob, db)
    assert torch.equal(got_ref, want_ref) and torch.equal(got_lc.cpu(), want_lc)
    # one object only, no depth
    one = torch.zeros(1, H, W)
    one[0, 5:25, 10:40] = 1
    r1, m1, ro1, d1 = TD.crop_rois(img.to(DEV), one.to(DEV), None)
    ro, mo, roo, do = O.crop_rois(img, one.clone(), None)
    assert torch.equal(roo, ro1.cpu()) and torch.equal(mo, m1.cpu()) and d1 is None
    a, _ = O.match_label_crop(one.clone(), (mo + 1).clone(), mo, roo, None)
    b, _ = TD.match_label_crop(one.clone(), (m1 + 1), m1, ro1, None)
    assert torch.equal(a, b)
    # nothing to filter / nothing to crop
    z = torch.zeros(1, H, W)
    assert torch.equal(TD.filter_labels_depth(z.to(DEV), xyz.to(DEV), 0.8).cpu(), z)
    assert TD.crop_rois(img.to(DEV), z.to(DEV), xyz.to(DEV))[0].shape[0] == 0


def test_multilabel_metrics_edge_cases():
    """Perfect prediction (every ratio 1), permuted ids, and a prediction with more objects than the ground truth."""
    from unseenobjectclustering_b200 import evaluation as EV
    _, gt = O.synthetic_clustered_features(72, 88, 8, 4, 0.05, 500)
    gt = gt.numpy().astype(np.float32)
    m = EV.multilabel_metrics(gt.copy(), gt)
    assert m['Objects F-measure'] == 1.0 and m['Boundary F-measure'] == 1.0 and m['obj_detected_075'] == m['obj_gt'] == 4
    perm = np.array([0, 9, 3, 17, 4], dtype=np.float32)[gt.astype(np.int64)]
    assert EV.multilabel_metrics(perm, gt) == m | {}
    pred = gt.copy()
    pred[:, 44:][pred[:, 44:] > 0] += 20                   # every object split in two along a vertical line
    got, want = EV.multilabel_metrics(pred, gt), O.multilabel_metrics(pred, gt)
    assert all(abs(float(got[k]) - float(want[k])) < 1e-12 for k in want) and got['obj_detected'] > got['obj_gt']


def test_two_stage_full_size_matches_reference_golden():
    """BASELINE config 3 at full size: 640x480 frame, 6 objects -> 6 crops of 224x224 through test_sample against the
    output of the unmodified reference's test_sample (tests/golden/full_cfg3.npz, oracle/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "full_cfg3.npz"))
    H, W, K = int(g["H"]), int(g["W"]), int(g["objects"])
    Xp, _ = O.exact_clustered_field(H, W, 64, K, 0.05, int(g["seed"]))
    assert O.field_crc32(Xp) == int(g["crc32"])
    fd = torch.from_numpy(Xp).view(1, 64, H, W).to(DEV)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=int(g["frame_seed"]))
    xyz[:, 2, :40, :] = 0
    crops = [torch.from_numpy(O.exact_clustered_field(224, 224, 64, 2, 0.05, int(s))[0]).view(1, 64, 224, 224).to(DEV)
             for s in g["crop_seeds"]]

    def net(i, l, dd):
        return fd

    def net_crop(i, l, dd):
        return torch.cat(crops[:i.shape[0]], 0)

    out_label, refined = TD.test_sample({'image_color': img, 'depth': xyz}, net, net_crop, [int(g["first_index"])],
                                        g["first_indices_crop"].tolist(), flags=_lib.FLAG_SYNC_CHECK)
    assert out_label.shape == (1, H, W) and refined is not None
    assert O.labels_equal_up_to_permutation(out_label.numpy(), g["out_label"])
    assert np.array_equal(out_label.numpy(), g["out_label"].astype(np.float32))
    assert np.array_equal(refined.numpy(), g["refined"].astype(np.float32))    # stage-2 ids are deterministic (1..K far to near)
    rgb_c, mask_c, rois, depth_c = TD.crop_rois(img.to(DEV), out_label, xyz.to(DEV))
    assert rgb_c.shape[0] == int(g["num_crops"]) and np.array_equal(rois.cpu().numpy(), g["rois"])
