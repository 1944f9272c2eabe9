"""GPU parity tests of the backbone: the tcgen05 implicit-GEMM convolution against torch's
convolution on bf16-rounded operands, and the whole two-branch ResNet34-8s against the golden
fixtures written by the unmodified reference (tolerance: 1e-3 cosine distance, BASELINE.json)."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import uoc_oracle as O
from conftest import GOLDEN
from unseenobjectclustering_b200 import _lib
from unseenobjectclustering_b200 import networks as NW

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CONV_CASES = [
    # Cin, Cout, k, stride, dil, H, W, N
    (64, 64, 3, 1, 1, 40, 56, 2),
    (64, 64, 3, 1, 1, 30, 44, 1),      # ragged tiles (H % 8, W % 16 != 0)
    (64, 128, 3, 2, 1, 40, 56, 2),     # stride 2 (TMA element strides)
    (64, 128, 1, 2, 1, 40, 56, 2),     # 1x1 stride-2 downsample
    (64, 128, 3, 2, 1, 31, 45, 1),     # stride 2, odd sizes
    (128, 256, 3, 1, 2, 20, 28, 2),    # dilation 2
    (256, 512, 3, 1, 4, 20, 28, 1),    # dilation 4
    (128, 256, 1, 1, 1, 20, 28, 1),
    (512, 64, 1, 1, 1, 20, 28, 2),     # fc shape
    (512, 512, 3, 1, 4, 60, 80, 1),    # layer4 at 640x480
    (64, 64, 3, 1, 1, 120, 160, 1),    # layer1 at 640x480
    (256, 256, 3, 1, 2, 28, 28, 3),    # layer3 on 224x224 crops, batch 3
    (128, 128, 3, 1, 1, 60, 80, 8),    # layer2, 4 frames x 2 branches: several stream-K segments per CTA pair
    (256, 512, 1, 1, 1, 60, 80, 8),    # 1x1 down-sample, 320 whole tiles dealt over 74 pairs (accumulator ring wraps)
    (256, 256, 3, 1, 2, 60, 80, 2),    # layer3 at 640x480: 38 tiles on 74 pairs, every tile split in two
    (128, 128, 3, 1, 1, 28, 28, 3),    # layer2 on 224x224 crops (weights-resident kernel: ragged 16 x 8 tiles, 2 K blocks)
    (64, 64, 3, 1, 1, 120, 160, 8),    # layer1, 4 frames: 9 tiles per CTA of the weights-resident kernel (accumulator ring wraps)
]


def _conv_call(x, w, bias, res, N, H, W, Cin, Cout, k, stride, dil, relu, flags):
    lib = _lib.load()
    pad = dil if k == 3 else 0
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=DEV)
    st = lib.uoc_conv2d_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(y), N, H, W, Cin, Cout, k,
                             stride, dil, relu, flags | _lib.FLAG_SYNC_CHECK, _lib.stream_ptr(torch.device(DEV)))
    _lib.check(st, "uoc_conv2d_bf16")
    return y


def _wres_case(case):
    Cin, Cout, k, stride, dil = case[:5]
    return k == 3 and stride == 1 and dil == 1 and Cin in (64, 128)


@pytest.mark.parametrize("variant", ["pair", "tc", "wres", "simt"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv_matches_torch(case, variant, knob):
    """The persistent CTA-pair stream-K kernel (conv_pair.cu), the one-tile-per-CTA kernel (conv_tc.cu), the
    weights-resident halo kernel of layers 1 / 2 (conv_wres.cu; only the shapes it takes) and the SIMT validation kernel,
    against torch's convolution on the same bf16 operands."""
    Cin, Cout, k, stride, dil, H, W, N = case
    if variant == "wres" and not _wres_case(case):
        pytest.skip("not a shape of the weights-resident kernel")
    flags = _lib.FLAG_CONV_SIMT if variant == "simt" else 0
    knob("conv_wres", 1 if variant == "wres" else 0)
    knob("conv_pair", 0 if variant == "tc" else 1)
    g = torch.Generator().manual_seed(Cin + Cout + k + H)
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(Cout, k * k, Cin, generator=g) * (1.0 / np.sqrt(k * k * Cin))).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g) * 0.1
    pad = dil if k == 3 else 0
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().view(Cout, k, k, Cin).permute(0, 3, 1, 2), bias, stride=stride,
                   padding=pad, dilation=dil)
    res = (torch.randn(ref.shape, generator=g) * 0.5).to(torch.bfloat16)
    ref_res = torch.relu(ref + res.float())
    xd, wd, bd = x.to(DEV), w.to(DEV), bias.to(DEV)
    y0 = _conv_call(xd, wd, bd, None, N, H, W, Cin, Cout, k, stride, dil, 0, flags).float().cpu().permute(0, 3, 1, 2)
    err0 = (y0 - ref).abs().max().item()
    assert err0 < 0.02 * max(1.0, ref.abs().max().item()), err0      # bf16 output rounding (2^-9 relative)
    resd = res.permute(0, 2, 3, 1).contiguous().to(DEV)
    y1 = _conv_call(xd, wd, bd, resd, N, H, W, Cin, Cout, k, stride, dil, 1, flags).float().cpu().permute(0, 3, 1, 2)
    err1 = (y1 - ref_res).abs().max().item()
    assert err1 < 0.02 * max(1.0, ref_res.abs().max().item()), err1
    assert (y1 >= 0).all()


def _golden_net(g, flags):
    sd = O.randomise_bn_(NW.random_state_dict(64, seed=int(g["weight_seed"])), int(g["weight_seed"]) + 1000)
    net = NW.seg_resnet34_8s_embedding(2, 64, sd).to(DEV)
    net.flags = flags | _lib.FLAG_SYNC_CHECK
    return net, sd


@pytest.mark.parametrize("flags", [0, _lib.FLAG_CONV_SIMT], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("name", ["backbone_a", "backbone_b"])
def test_backbone_matches_reference_golden(name, flags):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    H, W = int(g["H"]), int(g["W"])
    net, _ = _golden_net(g, flags)
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=int(g["frame_seed"]))
    f = net(img.to(DEV), None, xyz.to(DEV))
    assert f.shape == (1, 64, H, W) and f.dtype == torch.float32 and f.is_cuda
    ta = net.read_trunk(0, 1, H, W).cpu().numpy()
    tb = net.read_trunk(1, 1, H, W).cpu().numpy()
    for got, want in ((ta, g["trunk_rgb"]), (tb, g["trunk_depth"])):
        rel = np.abs(got - want).max() / np.abs(want).max()
        assert rel < 0.03, rel                                   # bf16 activations through 36 layers
    fc = f.cpu()
    assert (fc.norm(dim=1) - 1).abs().max() < 1e-4
    sub = fc[:, :, ::4, ::4].numpy()
    cosd = 1.0 - (sub * g["features_sub"]).sum(1)
    assert np.abs(cosd).max() < 1e-3, np.abs(cosd).max()          # BASELINE.json tolerance


def test_backbone_full_frame_vs_oracle_and_bf16_copy():
    from unseenobjectclustering_b200 import mean_shift as MS
    sd = O.randomise_bn_(NW.random_state_dict(64, seed=3), 1003)
    net = NW.seg_resnet34_8s_embedding(2, 64, sd).to(DEV)
    net.flags = _lib.FLAG_SYNC_CHECK
    img, xyz = O.synthetic_rgbd_frame(480, 640, seed=1)
    f = net(img.to(DEV), None, xyz.to(DEV))
    want = O.OracleSegNet(sd)(img, None, xyz)
    cosd = (1.0 - (f.cpu() * want).sum(1)).abs().max().item()
    assert cosd < 1e-3, cosd
    xb = MS._lookup_bf16(f)
    assert xb is not None and xb.shape == (1, 480 * 640, 64)
    back = xb.float().view(1, 480, 640, 64).permute(0, 3, 1, 2)
    assert (back - f).abs().max().item() < 1e-2
    # batch of 2 crops-like inputs vs two single calls: the stream-K convolution cuts the K range of a tile where the
    # launch's work divides evenly over the CTA pairs, i.e. at batch-dependent places, so the fp32 summation order (and
    # with it a few bf16 roundings per layer) depends on the batch: equal within a fraction of the embedding tolerance,
    # and a repeated call is bit-identical
    i2, x2 = O.synthetic_rgbd_frame(224, 224, seed=2)
    i3, x3 = O.synthetic_rgbd_frame(224, 224, seed=3)
    fb = net(torch.cat([i2, i3]).to(DEV), None, torch.cat([x2, x3]).to(DEV)).cpu()
    f2 = net(i2.to(DEV), None, x2.to(DEV)).cpu()
    f3 = net(i3.to(DEV), None, x3.to(DEV)).cpu()
    assert float((1.0 - (fb[0] * f2[0]).sum(0)).abs().max()) < 2e-4
    assert float((1.0 - (fb[1] * f3[0]).sum(0)).abs().max()) < 2e-4
    assert torch.equal(net(i2.to(DEV), None, x2.to(DEV)).cpu(), f2)


@pytest.mark.parametrize("H,W", [(100, 132), (90, 124), (75, 100), (61, 83), (33, 50)])
def test_backbone_takes_any_frame_size(H, W):
    """The reference takes every frame size (resnet_dilated.py:293,325: the trunk's stride-2 stages round as PyTorch's
    convolutions do, the head interpolates back to the input size).  Sizes that are not multiples of 8: ragged tiles in
    every layer, odd stem / max-pool / stride-2 extents; widths that are not multiples of 4 take the thread-staged stem and
    the one-pixel-per-thread head.  Same bars as the full frame: <= 1e-3 cosine distance, bf16 copy = the field."""
    sd = O.randomise_bn_(NW.random_state_dict(64, seed=4), 1004)
    net = NW.seg_resnet34_8s_embedding(2, 64, sd).to(DEV)
    net.flags = _lib.FLAG_SYNC_CHECK
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=H + W)
    f, xb = net.forward_ex(img.to(DEV), None, xyz.to(DEV), graph=False)
    want = O.OracleSegNet(sd)(img, None, xyz)
    assert f.shape == want.shape == (1, 64, H, W)
    assert bool(torch.isfinite(f).all())
    cosd = (1.0 - (f.cpu() * want).sum(1)).abs().max().item()
    assert cosd < 1e-3, cosd
    assert xb.shape == (1, H * W, 64)
    back = xb.float().view(1, H, W, 64).permute(0, 3, 1, 2)
    assert (back - f).abs().max().item() < 1e-2
    # and straight into the clustering with that bf16 copy: labels of a valid partition, every stage ran
    from unseenobjectclustering_b200 import mean_shift as MS
    import uoc_oracle_c as C
    labels, sel = MS.cluster_fields(f, 100, 20.0, 10, [H * W // 2], x_bf16=xb, flags=_lib.FLAG_SYNC_CHECK)[:2]
    sel_o, _ = C.select_seeds(f.cpu()[0].reshape(64, -1).numpy(), 100, H * W // 2)
    assert np.array_equal(sel.cpu().numpy()[0], sel_o)                   # the bf16 screen never changes an index
    assert labels.shape == (1, H * W) and int(labels.min()) >= 0


def test_graphed_forward_equals_eager_and_returns_fresh_tensors():
    """forward_ex replays a captured CUDA graph from the second call of a shape on: results are bit-identical to the
    launch-by-launch path, returned tensors are fresh (an earlier result is not overwritten by a later call), the bf16
    copy belongs to the returned tensor, and static_outputs=True hands out the graph's own buffers."""
    from unseenobjectclustering_b200 import mean_shift as MS
    net = NW.seg_resnet34_8s_embedding(2, 64, O.randomise_bn_(NW.random_state_dict(64, seed=5), 1005)).to(DEV)
    frames = [O.synthetic_rgbd_frame(96, 128, seed=s) for s in (11, 12, 13)]
    net.use_graphs = False
    eager = [net(i.to(DEV), None, x.to(DEV)).clone() for i, x in frames]
    net.use_graphs = True
    got = [net(i.to(DEV), None, x.to(DEV)) for i, x in frames]          # eager, capture + replay, replay
    assert len(net._graphs) == 1
    for a, b in zip(eager, got):
        assert torch.equal(a, b)
    assert got[1].data_ptr() != got[2].data_ptr()
    xb = MS._lookup_bf16(got[2])
    assert xb is not None and (xb.float().view(1, 96, 128, 64).permute(0, 3, 1, 2) - got[2]).abs().max().item() < 1e-2
    f_static, xb_static = net.forward_ex(frames[0][0].to(DEV), None, frames[0][1].to(DEV), static_outputs=True)
    assert torch.equal(f_static, eager[0])
    g = next(iter(net._graphs.values()))
    assert f_static.data_ptr() == g.out.data_ptr() and xb_static.data_ptr() == g.xb.data_ptr()
    assert torch.equal(got[1], eager[1])                                  # untouched by the later replays


def test_stem_without_tma_for_unaligned_inputs():
    """The stem fetches its input patches by TMA when the planes are 16-byte aligned; an input view that starts 4 bytes into
    its storage takes the thread-staged path: same result, bit for bit."""
    net = NW.seg_resnet34_8s_embedding(2, 64, O.randomise_bn_(NW.random_state_dict(64, seed=6), 1006)).to(DEV)
    img, xyz = O.synthetic_rgbd_frame(96, 128, seed=21)
    n = img.numel()
    ref, _ = net.forward_ex(img.to(DEV), None, xyz.to(DEV), graph=False)
    buf_i = torch.empty(n + 1, dtype=torch.float32, device=DEV)
    buf_x = torch.empty(n + 1, dtype=torch.float32, device=DEV)
    vi, vx = buf_i[1:].view(img.shape), buf_x[1:].view(xyz.shape)
    vi.copy_(img)
    vx.copy_(xyz)
    assert vi.data_ptr() % 16 == 4 and vi.is_contiguous()
    got, _ = net.forward_ex(vi, None, vx, graph=False)
    assert torch.equal(got, ref)


def test_module_drop_in_behaviour():
    net = NW.seg_resnet34_8s_embedding(2, 64, None).cuda(0)
    dp = torch.nn.DataParallel(net, device_ids=[0]).cuda(0)
    dp.eval()
    img, xyz = O.synthetic_rgbd_frame(64, 96, seed=0)
    f = dp(img.cuda(), None, xyz.cuda()).detach()
    assert f.shape == (1, 64, 64, 96)
    assert (f.norm(dim=1) - 1).abs().max() < 1e-4
    _ = f[0, 0::3]      # the indexing the reference's visualisation does


VARIANTS = ["variant_color", "variant_depth", "variant_early", "variant_cat", "variant_add_nonorm"]


@pytest.mark.parametrize("flags", [0, _lib.FLAG_CONV_SIMT], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("name", VARIANTS)
def test_network_variants_match_reference_golden(name, flags):
    """COLOR / DEPTH single trunk, early fusion (6-channel stem), cat fusion (128-channel field),
    EMBEDDING_NORMALIZATION off -- against fixtures written by the unmodified reference factories."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    it, ft, norm, cin = str(g["input_type"]), str(g["fusion_type"]), bool(g["normalize"]), int(g["in_channels"])
    H, W = int(g["H"]), int(g["W"])
    sd = O.randomise_bn_(NW.random_state_dict(64, seed=int(g["weight_seed"]), input_type=it, fusion_type=ft, in_channels=cin),
                         int(g["weight_seed"]) + 1000)
    factory = getattr(NW, str(g["factory"]))
    net = factory(2, 64, sd, input_type=it, fusion_type=ft, normalize=norm).to(DEV)
    net.flags = flags | _lib.FLAG_SYNC_CHECK
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=int(g["frame_seed"]))
    f = net(img.to(DEV) if it != "DEPTH" else None, None, xyz.to(DEV) if it != "COLOR" else None)
    C = 128 if (it == "RGBD" and ft == "cat") else 64
    assert f.shape == (1, C, H, W) and f.dtype == torch.float32 and f.is_cuda
    sub = f.cpu()[:, :, ::2, ::2].numpy()
    want = g["features_sub"]
    cosd = 1.0 - (sub * want).sum(1) / np.maximum(np.linalg.norm(sub, axis=1) * np.linalg.norm(want, axis=1), 1e-12)
    assert np.abs(cosd).max() < 1e-3, np.abs(cosd).max()          # BASELINE.json tolerance
    if norm:
        assert (f.norm(dim=1) - 1).abs().max() < 1e-4
    else:                                                          # un-normalised field: magnitudes must agree too
        rel = np.abs(sub - want).max() / np.abs(want).max()
        assert rel < 0.03, rel
    # the field feeds the clustering like the default variant's does (bf16 copy registered, d = C)
    from unseenobjectclustering_b200 import mean_shift as MS
    xb = MS._lookup_bf16(f)
    assert xb is not None and xb.shape == (1, H * W, C)


def test_variant_argument_errors():
    with pytest.raises(ValueError):
        NW.seg_resnet34_8s_embedding(2, 64, None, input_type="RGBD", fusion_type="early")     # needs the _early factory
    with pytest.raises(ValueError):
        NW.SEGNET_B200(64, None, input_type="LIDAR")
    net = NW.seg_resnet34_8s_embedding(2, 128, None, input_type="RGBD", fusion_type="cat").to(DEV)
    img, xyz = O.synthetic_rgbd_frame(32, 32, seed=0)
    with pytest.raises(_lib.UocError):                               # cat fusion is built for num_units = 64 only
        net(img.to(DEV), None, xyz.to(DEV))
    net = NW.seg_resnet34_8s_embedding(2, 64, None, input_type="COLOR").to(DEV)
    with pytest.raises(_lib.UocError):
        net(None, None, xyz.to(DEV))
