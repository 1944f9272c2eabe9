"""GPU parity tests of the clustering path, through the C ABI (ctypes), against the oracles.
Bit-exact for index / label work against the canonical-order C oracle; tolerance (stated per
test) for the floating-point mean-shift loop; golden fixtures = outputs of the unmodified reference."""
import glob
import os

import numpy as np
import pytest
import torch

import uoc_oracle as O
import uoc_oracle_c as C
from conftest import GOLDEN
from unseenobjectclustering_b200 import _lib
from unseenobjectclustering_b200 import mean_shift as MS
from unseenobjectclustering_b200 import test_dataset as TD

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _field(H, W, d, K, noise, seed):
    feats, gt = O.synthetic_clustered_features(H, W, d, K, noise, seed)
    return feats, gt


def _fps_mode(knob, mode):
    """tc: tcgen05 screen (fps_tc.cu, default: tiles in tensor memory first); tc_smem: same with every tile in shared memory;
    stream: the streaming bf16 screen for fields that do not fit on chip (fps5_kernel), forced on small fields;
    fp32: no screen (fps2_kernel).  Returns bf16_screen."""
    knob("fps_tc", 1 if mode in ("tc", "tc_smem", "stream") else 0)
    knob("fps_stream", 1 if mode == "stream" else 0)
    if mode == "tc_smem":
        knob("fps_tmem_tiles", 0)
    return mode != "fp32"


@pytest.mark.parametrize("mode", ["tc", "tc_smem", "stream", "fp32"])
@pytest.mark.parametrize("H,W,d,m", [(32, 48, 64, 100), (37, 41, 64, 100), (30, 44, 128, 100), (24, 24, 64, 40),
                                     (60, 80, 32, 17), (120, 160, 64, 100)])
def test_select_seeds_bit_exact(H, W, d, m, mode, knob):
    screen = _fps_mode(knob, mode)
    feats, _ = _field(H, W, d, 4, 0.05, seed=H * 7 + d)
    Xp = feats[0].reshape(d, -1).numpy()
    first = (H * W) // 3
    sel_o, seeds_o = C.select_seeds(Xp, m, first)
    X = feats.to(DEV)[0].view(d, -1).t()
    seeds, sel = MS.select_smart_seeds(X, m, return_selected_indices=True, first_index=first, bf16_screen=screen)
    assert np.array_equal(sel.numpy(), sel_o)
    assert np.array_equal(seeds.cpu().numpy(), seeds_o)
    assert sel[0] == first and len(set(sel.tolist())) == m


@pytest.mark.parametrize("mode", ["tc", "stream"])
@pytest.mark.parametrize("case", ["isotropic", "scaled", "two_homes", "full_frame", "full_frame_128"])
def test_select_seeds_bf16_screen_stress(case, mode, knob):
    """The bf16 screening pass of the seed selection (fps_tc.cu) must never change an index:
    isotropic  - no cluster structure: the screen rejects little, nearly every point takes the fp32 path;
    scaled     - rows of norm 3 (the error bound of the screen scales with |x| |s|);
    two_homes  - one tile in tensor memory, the second in shared memory (small field, both operand homes);
    full_frame - 480x640x64: 17 tiles per CTA (7 in tensor memory + 10 in shared memory for fps_tc);
    full_frame_128 - 240x320x128: the two-block (d = 128) operand layouts at several tiles per CTA.
    mode stream: the same inputs through the streaming screen (fps5_kernel)."""
    _fps_mode(knob, mode)
    if case == "isotropic":
        g = torch.Generator().manual_seed(5)
        feats = torch.nn.functional.normalize(torch.randn(1, 64, 48, 64, generator=g), dim=1)
    elif case == "scaled":
        feats = _field(40, 52, 64, 4, 0.1, seed=12)[0] * 3.0
    elif case == "two_homes":
        knob("fps_tmem_tiles", 1)
        feats = _field(120, 160, 64, 5, 0.1, seed=13)[0]
    elif case == "full_frame":
        feats = _field(480, 640, 64, 6, 0.1, seed=14)[0]
    else:
        feats = _field(240, 320, 128, 6, 0.1, seed=15)[0]
    d = feats.shape[1]
    m = 100
    Xp = feats[0].reshape(d, -1).numpy()
    first = Xp.shape[1] // 5
    sel_o, seeds_o = C.select_seeds(Xp, m, first)
    X = feats.to(DEV)[0].view(d, -1).t()
    seeds, sel = MS.select_smart_seeds(X, m, return_selected_indices=True, first_index=first, bf16_screen=True)
    assert np.array_equal(sel.numpy(), sel_o)
    assert np.array_equal(seeds.cpu().numpy(), seeds_o)


@pytest.mark.parametrize("flags,tol", [(0, 5e-5), (_lib.FLAG_LOOP_SIMT, 2e-6)], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("H,W,d,m", [(32, 48, 64, 100), (37, 41, 64, 50), (30, 44, 128, 100), (96, 128, 64, 100),
                                     (40, 56, 64, 7), (40, 56, 64, 64), (40, 56, 64, 65), (40, 56, 64, 128)])
def test_hill_climb_vs_double_oracle(H, W, d, m, flags, tol):
    """Cosine distance between converged seeds and the double-precision oracle: tcgen05 loop (bf16 operands, fp32
    accumulate) <= 5e-5, fp32 SIMT loop <= 2e-6.  m = 7 / 64 / 65 / 128: few seeds, and the full 128 seed rows."""
    feats, _ = _field(H, W, d, 4, 0.05, seed=3 + H)
    Xp = feats[0].reshape(d, -1).numpy()
    _, seeds = C.select_seeds(Xp, m, 5)
    Zo = C.hill_climb(Xp, seeds, 20.0, 10)
    X = feats.to(DEV)[0].view(d, -1).t()
    Z = MS.seed_hill_climbing_ball(X, torch.from_numpy(seeds).to(DEV), 20.0, 10, flags=flags).cpu().numpy()
    assert np.isfinite(Z).all()
    cosd = 1.0 - (Z * Zo).sum(1)
    assert np.abs(cosd).max() < tol, np.abs(cosd).max()
    assert np.abs(np.linalg.norm(Z, axis=1) - 1).max() < 1e-5


def test_hill_climb_one_iteration_matches_torch():
    feats, _ = _field(64, 64, 64, 3, 0.05, seed=9)
    X = feats[0].view(64, -1).t()
    seeds, _ = O.select_seeds(X, 100, 7)
    Zt = O.hill_climb(X, seeds, 20.0, 1)
    Z = MS.seed_hill_climbing_ball(X.to(DEV), seeds.to(DEV), 20.0, 1).cpu()
    assert (1 - (Z * Zt).sum(1)).abs().max() < 2e-5


@pytest.mark.parametrize("m,d", [(100, 64), (100, 128), (37, 64), (128, 32)])
def test_label_seeds_bit_exact(m, d):
    rng = np.random.RandomState(m + d)
    for trial in range(6):
        base = rng.randn(6, d).astype(np.float32)
        Z = base[rng.randint(0, 6, m)] + (0.02 + 0.03 * trial) * rng.randn(m, d).astype(np.float32)
        Z /= np.linalg.norm(Z, axis=1, keepdims=True)
        lo, uo = C.label_seeds(Z, 0.04)
        lg, ug = MS.connected_components(torch.from_numpy(Z).to(DEV), 0.04, return_num_unique=True)
        assert np.array_equal(lg.numpy(), lo), trial
        assert ug == uo


@pytest.mark.parametrize("H,W,d,m", [(32, 48, 64, 100), (37, 41, 64, 33), (30, 44, 128, 100), (20, 30, 32, 10)])
def test_assign_bit_exact(H, W, d, m):
    feats, _ = _field(H, W, d, 4, 0.05, seed=21 + d)
    Xp = feats[0].reshape(d, -1).numpy()
    _, seeds = C.select_seeds(Xp, m, 1)
    Z = C.hill_climb(Xp, seeds, 20.0, 3)
    sl, uniq = C.label_seeds(Z, 0.04)
    lo = C.assign(Xp, Z, sl, uniq)
    X = feats.to(DEV)[0].view(d, -1).t()
    lg = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(sl), uniq)
    assert np.array_equal(lg.numpy(), lo)
    # gapped labels exercise the range(len(unique)) histogram quirk
    sl2 = (sl * 2).astype(np.int32)
    lo2 = C.assign(Xp, Z, sl2, uniq)
    lg2 = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(sl2), uniq)
    assert np.array_equal(lg2.numpy(), lo2)
    if d in (64, 128):
        # tensor-core pass with exactness certificate + fp32 fix-up: identical labels
        lt = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(sl), uniq, use_tensor_cores=True)
        assert np.array_equal(lt.numpy(), lo)
        lt2 = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(sl2), uniq, use_tensor_cores=True)
        assert np.array_equal(lt2.numpy(), lo2)


@pytest.mark.parametrize("noise", [0.05, 0.2, 0.5])
def test_assign_tensor_core_certificate_on_marginless_input(noise):
    """Every seed its own label + heavy noise: most points sit near a decision border, so the certificate fails
    often and the fp32 fix-up path does the work; the result must still equal the canonical arg-min."""
    H, W, d, m = 60, 84, 64, 100
    feats, _ = _field(H, W, d, 5, noise, seed=int(noise * 100))
    Xp = feats[0].reshape(d, -1).numpy()
    _, seeds = C.select_seeds(Xp, m, 3)
    Z = C.hill_climb(Xp, seeds, 20.0, 2)
    sl = np.arange(m, dtype=np.int32)                 # all different labels: arg-min ties/borders everywhere
    lo = C.assign(Xp, Z, sl, m)
    X = feats.to(DEV)[0].view(d, -1).t()
    lt = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(sl), m, use_tensor_cores=True)
    assert np.array_equal(lt.numpy(), lo)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "cluster_*.npz")))


@pytest.mark.parametrize("flags", [0, _lib.FLAG_LOOP_SIMT])
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_full_clustering_matches_reference_golden(path, flags):
    g = np.load(path)
    feats = torch.from_numpy(g["features"]).to(DEV)
    m = int(g["num_seeds"])
    labels, sel, Z, sl = MS.cluster_fields(feats, m, first_indices=[int(g["first_index"])], flags=flags | _lib.FLAG_SYNC_CHECK,
                                           return_seeds=True)
    assert np.array_equal(sel[0].cpu().numpy(), g["selected"])
    cosd = 1.0 - (Z[0].cpu().numpy() * g["Z"]).sum(1)
    assert np.abs(cosd).max() < 2e-5
    assert O.labels_equal_up_to_permutation(labels[0].cpu().numpy(), g["labels"])
    # on these inputs even the ids agree
    assert np.array_equal(sl[0].cpu().numpy(), g["seed_labels"])
    assert np.array_equal(labels[0].cpu().numpy(), g["labels"])


def test_reference_api_surface_types():
    g = np.load(FIXTURES[0])
    feats = torch.from_numpy(g["features"]).to(DEV)
    np.random.seed(3)
    expect_first = np.random.randint(0, feats.shape[2] * feats.shape[3])
    np.random.seed(3)
    out_label, picked = TD.clustering_features(feats, num_seeds=int(g["num_seeds"]))
    assert out_label.dtype == torch.float32 and out_label.device.type == "cpu" and out_label.shape == (1, int(g["H"]), int(g["W"]))
    assert len(picked) == 1 and picked[0].dtype == torch.int64 and picked[0].device.type == "cpu"
    assert int(picked[0][0]) == expect_first == int(g["first_index"])
    assert np.array_equal(out_label.numpy().ravel(), g["labels"].astype(np.float32))
    X = feats[0].view(feats.shape[1], -1).t()
    lab, sel = MS.mean_shift_smart_init(X, kappa=20, num_seeds=int(g["num_seeds"]), max_iters=10, metric='cosine',
                                        first_index=int(g["first_index"]))
    assert lab.dtype == torch.int64 and lab.device.type == "cpu" and np.array_equal(lab.numpy(), g["labels"])
    # pixel-major input (non planar strides) is accepted too
    lab2, _ = MS.mean_shift_smart_init(X.contiguous(), kappa=20, num_seeds=int(g["num_seeds"]), max_iters=10,
                                       first_index=int(g["first_index"]))
    assert torch.equal(lab, lab2)


def test_batched_call_equals_per_item_calls():
    fields = [_field(40, 56, 64, 3, 0.05, seed=40 + k)[0] for k in range(5)]
    batch = torch.cat(fields, 0).to(DEV)
    firsts = [11, 222, 333, 444, 555]
    lb, sb = MS.cluster_fields(batch, 100, first_indices=firsts, flags=_lib.FLAG_SYNC_CHECK)
    for k in range(5):
        l1, s1 = MS.cluster_fields(fields[k].to(DEV), 100, first_indices=[firsts[k]], flags=_lib.FLAG_SYNC_CHECK)
        assert torch.equal(lb[k], l1[0]) and torch.equal(sb[k], s1[0])


def test_full_resolution_properties():
    """640x480x64 (BASELINE config 2): ground-truth recovery up to permutation, determinism,
    distinct seeds, label 0 = largest cluster."""
    feats, gt = _field(480, 640, 64, 6, 0.05, seed=0)
    f = feats.to(DEV)
    l1, s1 = MS.cluster_fields(f, 100, first_indices=[71530], flags=_lib.FLAG_SYNC_CHECK)
    l2, s2 = MS.cluster_fields(f, 100, first_indices=[71530], flags=_lib.FLAG_SYNC_CHECK)
    assert torch.equal(l1, l2) and torch.equal(s1, s2)
    lab = l1[0].cpu().numpy()
    assert O.labels_equal_up_to_permutation(lab, gt.numpy().ravel())
    counts = np.bincount(lab)
    assert counts.argmax() == 0
    sel = s1[0].cpu().numpy()
    assert sel[0] == 71530 and len(set(sel.tolist())) == 100
    # first 8 seeds agree with the canonical oracle (cheap partial check at full size)
    sel_o, _ = C.select_seeds(feats[0].reshape(64, -1).numpy(), 8, 71530)
    assert np.array_equal(sel[:8], sel_o)


def test_high_res_128d_30_iters():
    """BASELINE config 5 shape class (d = 128, 30 iterations) at reduced resolution."""
    feats, gt = _field(180, 240, 128, 8, 0.04, seed=5)
    l, s = MS.cluster_fields(feats.to(DEV), 100, max_iters=30, first_indices=[5], flags=_lib.FLAG_SYNC_CHECK)
    assert O.labels_equal_up_to_permutation(l[0].cpu().numpy(), gt.numpy().ravel())


def test_config5_full_size_960x720_128d_30_iters():
    """BASELINE config 5 at full size (n = 691 200, d = 128, 30 updates): 354 MB field, > L2."""
    feats, gt = _field(720, 960, 128, 12, 0.04, seed=55)
    f = feats.to(DEV)
    l, s = MS.cluster_fields(f, 100, max_iters=30, first_indices=[123456], flags=_lib.FLAG_SYNC_CHECK)
    assert O.labels_equal_up_to_permutation(l[0].cpu().numpy(), gt.numpy().ravel())
    sel = s[0].cpu().numpy()
    assert sel[0] == 123456 and len(set(sel.tolist())) == 100
    # the field does not fit on chip: the streaming bf16 screen (fps5_kernel) picks the seeds; same indices as the oracle
    sel_o, _ = C.select_seeds(feats[0].reshape(128, -1).numpy(), 24, 123456)
    assert np.array_equal(sel[:24], sel_o)


def test_bad_arguments_fail_loudly():
    f = torch.zeros(1, 64, 8, 8, device=DEV)
    with pytest.raises(_lib.UocError):
        MS.cluster_fields(f, 200)                      # > UOC_MAX_SEEDS
    with pytest.raises(_lib.UocError):
        MS.cluster_fields(f, 10, first_indices=[64])   # first seed out of range
    with pytest.raises(_lib.UocError):
        MS.cluster_fields(f.cpu(), 10)


# ----------------------------------------------------------------------------------------------
# metric='euclidean' (lib/utils/mean_shift.py:21-24,58-60,101-105,159-160,207-209)
# ----------------------------------------------------------------------------------------------
EUCLID_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "euclid_*.npz")))


@pytest.mark.parametrize("path", EUCLID_FIXTURES, ids=[os.path.basename(p) for p in EUCLID_FIXTURES])
def test_euclidean_matches_reference_golden(path):
    """Whole euclidean clustering against fixtures written by the unmodified reference (euclid_b is not unit norm):
    selected indices, seed labels and pixel labels identical, converged seeds within 2e-5."""
    g = np.load(path)
    d, m = int(g["d"]), int(g["num_seeds"])
    feats = torch.from_numpy(g["features"]).to(DEV)
    labels, sel, Z, sl = MS.cluster_fields(feats, m, 20.0, 10, [int(g["first_index"])], flags=_lib.FLAG_SYNC_CHECK,
                                           return_seeds=True, metric="euclidean")
    assert np.array_equal(sel[0].cpu().numpy(), g["selected"])
    assert np.abs(Z[0].cpu().numpy() - g["Z"]).max() < 2e-5
    assert np.array_equal(sl[0].cpu().numpy(), g["seed_labels"])
    assert np.array_equal(labels[0].cpu().numpy(), g["labels"])
    # the reference-named entry point returns the reference's types
    X = feats[0].view(d, -1).t()
    l2, s2 = MS.mean_shift_smart_init(X, 20.0, m, 10, metric="euclidean", first_index=int(g["first_index"]))
    assert l2.dtype == torch.int64 and not l2.is_cuda and np.array_equal(l2.numpy(), g["labels"])
    assert np.array_equal(s2.numpy(), g["selected"])


@pytest.mark.parametrize("H,W,d,m,scale", [(32, 48, 64, 100, 1.0), (37, 41, 64, 50, 2.0), (30, 44, 128, 100, 1.0),
                                           (60, 80, 32, 17, 0.7), (33, 35, 20, 9, 1.0)])
def test_euclidean_stages_bit_exact_vs_c_oracle(H, W, d, m, scale):
    """Every discrete decision of the euclidean path against the canonical-order C oracle (bit-exact), stage by stage;
    the fp32 loop against the double-precision oracle (<= 1e-5 absolute)."""
    feats, _ = _field(H, W, d, 4, 0.03, seed=H + d)
    feats = feats * scale
    Xp = feats[0].reshape(d, -1).numpy()
    first = (H * W) // 3
    sel_o, seeds_o = C.select_seeds(Xp, m, first, metric="euclidean")
    X = feats.to(DEV)[0].view(d, -1).t()
    seeds, sel = MS.select_smart_seeds(X, m, return_selected_indices=True, first_index=first, metric="euclidean")
    assert np.array_equal(sel.numpy(), sel_o)
    assert np.array_equal(seeds.cpu().numpy(), seeds_o)
    Zo = C.hill_climb(Xp, seeds_o, 20.0, 10, metric="euclidean")
    Z = MS.seed_hill_climbing_ball(X, seeds, 20.0, 10, metric="euclidean").cpu().numpy()
    assert np.isfinite(Z).all() and np.abs(Z - Zo).max() < 1e-5 * max(1.0, scale), np.abs(Z - Zo).max()
    lo, uo = C.label_seeds(Z, 0.04, metric="euclidean")
    lg, ug = MS.connected_components(torch.from_numpy(Z).to(DEV), 0.04, metric="euclidean", return_num_unique=True)
    assert np.array_equal(lg.numpy(), lo) and ug == uo
    want = C.assign(Xp, Z, lo, uo, metric="euclidean")
    got = MS.assign_labels(X, torch.from_numpy(Z).to(DEV), torch.from_numpy(lo), uo, metric="euclidean")
    assert np.array_equal(got.numpy(), want)


def test_euclidean_weight_sum_clamp():
    """Isolated seeds (sum of weights < 1) are divided by 1, not by the sum (mean_shift.py:103-104)."""
    g = torch.Generator().manual_seed(3)
    X = torch.randn(64, 2000, generator=g) * 2.0                 # far-apart points: every weight but the self-weight ~ 0
    Xp = X.numpy()
    seeds = (X[:, :10].t().contiguous() + 0.05)                   # one neighbour at distance 0.4: sum of weights ~ 0.04 < 1
    Zo = C.hill_climb(Xp, seeds.numpy(), 20.0, 1, metric="euclidean")
    Z = MS.seed_hill_climbing_ball(X.t().to(DEV), seeds.to(DEV), 20.0, 1, metric="euclidean").cpu().numpy()
    assert 1e-3 < np.abs(Zo).max() < 1.0 and np.abs(Z - Zo).max() < 1e-6


def test_euclidean_clustering_features_and_batch():
    """clustering_features(metric='euclidean') on a batch equals the per-item calls and the torch oracle."""
    fa, _ = _field(32, 40, 64, 3, 0.02, seed=5)
    fb, _ = _field(32, 40, 64, 4, 0.02, seed=6)
    feats = torch.cat([fa, fb * 1.3], 0)
    out, sel = TD.clustering_features(feats.to(DEV), 60, [11, 700], metric="euclidean")
    want, wsel = O.clustering_features(feats, 60, [11, 700], metric="euclidean")
    assert out.dtype == torch.float32 and not out.is_cuda
    for j in range(2):
        assert torch.equal(sel[j], wsel[j])
        assert O.labels_equal_up_to_permutation(out[j].numpy().ravel(), want[j].numpy().ravel())
    with pytest.raises(ValueError):
        TD.clustering_features(feats.to(DEV), 60, [11, 700], metric="manhattan")


# ---------------------------------------------------------------------------------------------
# BASELINE-size fixtures written by the unmodified reference (oracle/make_golden.py gen_full)
# ---------------------------------------------------------------------------------------------
def _full_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    Xp, gt = O.exact_clustered_field(int(g["H"]), int(g["W"]), int(g["d"]), int(g["objects"]), float(g["noise"]), int(g["seed"]))
    assert O.field_crc32(Xp) == int(g["crc32"]), "exact_clustered_field is not bit-reproducible on this platform"
    feats = torch.from_numpy(Xp).view(1, int(g["d"]), int(g["H"]), int(g["W"]))
    return g, feats, gt


@pytest.mark.parametrize("name", ["full_cfg2", "full_cfg5"])
def test_full_size_clustering_matches_reference_golden(name):
    """640x480x64 / 10 updates (config 2) and 960x720x128 / 30 updates (config 5): ALL 100 selected indices, the converged
    seeds (<= 5e-5 cosine distance: bf16 operands in the tcgen05 loop) and all pixel labels against the reference's output."""
    g, feats, gt = _full_case(name)
    m = int(g["num_seeds"])
    labels, sel, Z, sl = MS.cluster_fields(feats.to(DEV), m, max_iters=int(g["max_iters"]), first_indices=[int(g["first_index"])],
                                           flags=_lib.FLAG_SYNC_CHECK, return_seeds=True)
    assert np.array_equal(sel[0].cpu().numpy(), g["selected"])
    cosd = 1.0 - (Z[0].cpu().numpy().astype(np.float64) * g["Z"].astype(np.float64)).sum(1)
    assert np.abs(cosd).max() < 5e-5, np.abs(cosd).max()
    lab = labels[0].cpu().numpy()
    assert O.labels_equal_up_to_permutation(lab, g["labels"])
    assert O.labels_equal_up_to_permutation(lab, gt.ravel())
    assert np.array_equal(sl[0].cpu().numpy(), g["seed_labels"])           # on these inputs even the ids agree
    assert np.array_equal(lab, g["labels"].astype(np.int32))


def test_full_size_loop_vs_double_oracle():
    """The mean-shift loop at 640x480x64 against the double-precision C oracle (same seeds in, 10 updates)."""
    g, feats, _ = _full_case("full_cfg2")
    Xp = feats[0].reshape(64, -1).numpy()
    X = feats.to(DEV)[0].view(64, -1).t()
    _, seeds_o = C.select_seeds(Xp, 100, int(g["first_index"]))
    Zo = C.hill_climb(Xp, seeds_o, 20.0, 10)
    Z = MS.seed_hill_climbing_ball(X, torch.from_numpy(seeds_o).to(DEV), 20.0, 10)
    cosd = 1.0 - (Z.cpu().numpy().astype(np.float64) * Zo.astype(np.float64)).sum(1)
    assert np.abs(cosd).max() < 5e-5, np.abs(cosd).max()


@pytest.mark.parametrize("name", ["init_a", "init_b"])
def test_select_seeds_with_init_seeds(name):
    """select_smart_seeds(init_seeds=, num_init_seeds=) (lib/utils/mean_shift.py:144-149, :164-169): indices bit-exact against
    the fixture written by the unmodified reference and against the C oracle; init_seeds is filled in place and returned."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, m, k, metric = int(g["d"]), int(g["num_seeds"]), int(g["num_init"]), str(g["metric"])
    feats, _ = O.synthetic_clustered_features(int(g["H"]), int(g["W"]), d, int(g["objects"]), float(g["noise"]), int(g["seed"]))
    feats = feats * float(g["scale"])
    X = feats.to(DEV)[0].view(d, -1).t()
    init = torch.zeros((m, d), device=DEV)
    init[:k] = torch.from_numpy(g["given"]).to(DEV)
    seeds, sel = MS.select_smart_seeds(X, m, return_selected_indices=True, init_seeds=init, num_init_seeds=k, metric=metric)
    assert seeds is init
    assert np.array_equal(sel.numpy(), g["selected"])
    assert np.array_equal(seeds.cpu().numpy(), g["seeds"])
    sel_c, seeds_c = C.select_seeds_init(feats[0].reshape(d, -1).numpy(), m, g["given"], metric)
    assert np.array_equal(sel.numpy(), sel_c) and np.array_equal(seeds.cpu().numpy(), seeds_c)


def test_bf16_side_channel_cannot_go_stale():
    """the bf16 copy of a backbone output must never be served for a different tensor that happens to get the
    same address.  Free a backbone output, allocate a same-shape foreign field, cluster it: the lookup must miss and the
    labels must be those of the foreign field."""
    from unseenobjectclustering_b200 import networks as NW
    MS.clear_bf16_registry()
    H, W = 64, 96
    net = NW.seg_resnet34_8s_embedding(2, 64, NW.random_state_dict(64, seed=0)).to(DEV)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=1)
    f = net(img.to(DEV), None, xyz.to(DEV))
    assert MS._lookup_bf16(f) is not None and MS._lookup_bf16(f.detach()) is not None
    addr = f.data_ptr()
    del f
    torch.cuda.synchronize()
    foreign, gt = O.synthetic_clustered_features(H, W, 64, 4, 0.05, seed=9)
    cands = [foreign.to(DEV) for _ in range(4)]
    assert all(c.data_ptr() != addr for c in cands), "the registered tensor is kept alive: its address cannot be recycled"
    for c in cands:
        assert MS._lookup_bf16(c) is None
    lab, _ = MS.cluster_fields(cands[0], 100, first_indices=[7], flags=_lib.FLAG_SYNC_CHECK)
    assert O.labels_equal_up_to_permutation(lab[0].cpu().numpy(), gt.numpy().ravel())
    # after eviction the address may be recycled -- and then there is no entry to hit
    MS.clear_bf16_registry()
    torch.cuda.empty_cache()
    again = [foreign.to(DEV) for _ in range(8)]
    assert all(MS._lookup_bf16(c) is None for c in again)
    lab2, _ = MS.cluster_fields(again[0], 100, first_indices=[7], flags=_lib.FLAG_SYNC_CHECK)
    assert torch.equal(lab, lab2)
    # in-place modification of a registered field invalidates its copy
    f = net(img.to(DEV), None, xyz.to(DEV))
    f.copy_(cands[0])
    assert MS._lookup_bf16(f) is None
    lab3, _ = MS.cluster_fields(f, 100, first_indices=[7], flags=_lib.FLAG_SYNC_CHECK)
    assert torch.equal(lab, lab3)
