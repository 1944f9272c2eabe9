"""N>1 host logic on CPU: world_size-2 gloo processes shard frames, label each with a stand-in
clusterer, all-gather the label maps; every rank must hold the same result as a single process."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_network(img, label, depth):
    return img                                  # "features" = the image itself


def _fake_cluster(feats, firsts):
    # deterministic integer map that depends on the frame content and on the pre-drawn first index
    return ((feats[0, 0] * 10).floor().to(torch.int32) + int(firsts[0]) % 7).view(1, -1)


def _worker(rank, world, port, frames, depths, firsts, out_dir):
    sys.path.insert(0, ROOT)
    from unseenobjectclustering_b200 import distributed as UD
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = UD.segment_frames(frames, depths, _fake_network, firsts, _fake_cluster, rank, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), res.numpy())
    dist.destroy_process_group()


def test_shard_range_partitions_every_frame_once():
    from unseenobjectclustering_b200 import distributed as UD
    for F in (1, 5, 8, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                b, e = UD.shard_range(F, world, r)
                seen.extend(range(b, e))
            assert seen == list(range(F))
    assert UD.draw_first_indices(5, 100, 3) == UD.draw_first_indices(7, 100, 3)[:5]
    np.random.seed(3)
    assert UD.draw_first_indices(3, 307200, 3)[0] == np.random.randint(0, 307200) == 71530


def test_two_rank_gloo_matches_single_process(tmp_path):
    from unseenobjectclustering_b200 import distributed as UD
    g = torch.Generator().manual_seed(0)
    F = 5                                                # odd: exercises the tail-rank padding
    frames = torch.rand(F, 3, 12, 16, generator=g)
    depths = torch.rand(F, 3, 12, 16, generator=g)
    firsts = UD.draw_first_indices(F, 12 * 16, 3)
    single = UD.segment_frames(frames, depths, _fake_network, firsts, _fake_cluster, 0, 1)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, frames, depths, firsts, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert np.array_equal(got, single.numpy())


def _verify_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from unseenobjectclustering_b200 import distributed as UD
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(10 + rank)
    mine = torch.randint(0, 9, (2, 12 * 16), generator=g, dtype=torch.uint8)       # two frames per rank
    gathered = UD.gather_labels(mine).view(world, -1)
    good = UD.verify_gathered_labels(mine.view(-1), gathered)
    broken = gathered.clone()
    if rank == 1:                                                                  # one rank, one pixel: every rank must say no
        broken[0, 7] += 1
    bad = UD.verify_gathered_labels(mine.view(-1), broken)
    swapped = gathered.flip(0)                                                     # same sums overall, wrong rank order
    swp = UD.verify_gathered_labels(mine.view(-1), swapped)
    np.save(os.path.join(out_dir, "verdict%d.npy" % rank), np.array([good, bad, swp]))
    dist.destroy_process_group()


def test_label_checksums_see_value_and_position():
    from unseenobjectclustering_b200 import distributed as UD
    a = torch.tensor([[0, 1, 2, 3]], dtype=torch.uint8)
    assert UD.label_checksums(a).tolist() == [[6, 0 * 1 + 1 * 2 + 2 * 3 + 3 * 4]]
    assert UD.label_checksums(a.flip(1))[0, 0] == 6 and UD.label_checksums(a.flip(1))[0, 1] != 20
    assert UD.label_checksums(torch.zeros((2, 0), dtype=torch.uint8)).tolist() == [[0, 0], [0, 0]]


def test_two_rank_gloo_gather_verification(tmp_path):
    port = _free_port()
    mp.spawn(_verify_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(os.path.join(str(tmp_path), "verdict%d.npy" % r)).tolist() == [True, False, False]
