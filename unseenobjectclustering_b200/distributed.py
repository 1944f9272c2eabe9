"""Frame sharding across the GPUs of one box: frames are independent units, so the
path shards with no data-path collective; the only exchange is ONE all-gather of the per-frame label
maps.  One process per GPU (torchrun), NCCL over NVLink on the GPU box, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(num_frames, world_size, rank):
    """Static block partition: frame f -> rank f // ceil(F / world).  Returns (begin, end)."""
    per = (num_frames + world_size - 1) // world_size
    b = min(rank * per, num_frames)
    return b, min(b + per, num_frames)


def draw_first_indices(num_frames, n, seed):
    """The reference consumes one np.random.randint(0, n) per frame from a global stream
    (lib/utils/mean_shift.py:155, seeded with cfg.RNG_SEED in tools/test_net.py:76).  Every rank
    draws the whole stream so that frame f gets the same first seed as in the sequential reference."""
    rs = np.random.RandomState(seed)
    return [int(rs.randint(0, n)) for _ in range(num_frames)]


def gather_labels(local_labels, num_frames=None, group=None):
    """all-gather per-rank label maps [F_local, ...] (equal F_local on every rank; pad the tail rank)
    into [world * F_local, ...] on every rank; trims to num_frames if given."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_labels if num_frames is None else local_labels[:num_frames]
    world = dist.get_world_size(group)
    out = torch.empty((world * local_labels.shape[0],) + tuple(local_labels.shape[1:]), dtype=local_labels.dtype,
                      device=local_labels.device)
    if local_labels.is_cuda:
        dist.all_gather_into_tensor(out, local_labels.contiguous(), group=group)
    else:
        parts = list(out.chunk(world, 0))
        dist.all_gather(parts, local_labels.contiguous(), group=group)
    return out if num_frames is None else out[:num_frames]


def _all_gather_rows(row, group=None):
    """[L] per rank -> [world, L] on every rank (NCCL: one all_gather_into_tensor; gloo: all_gather of the chunks)."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(row.shape), dtype=row.dtype, device=row.device)
    if row.is_cuda:
        dist.all_gather_into_tensor(out, row.contiguous(), group=group)
    else:
        dist.all_gather(list(out.unbind(0)), row.contiguous(), group=group)
    return out


def label_checksums(labels_u8):
    """[R, L] uint8 label maps (one flattened row per rank) -> [R, 2] int64: plain sum and position-weighted sum."""
    v = labels_u8.reshape(labels_u8.shape[0], -1).to(torch.int64)
    w = torch.arange(1, v.shape[1] + 1, device=v.device, dtype=torch.int64) % 65521
    return torch.stack([v.sum(1), (v * w[None]).sum(1)], 1)


def verify_gathered_labels(local_u8, gathered_u8, group=None):
    """True on every rank iff every rank's uint8 label maps arrived intact in `gathered_u8` [world, L] on every rank:
    the checksums of the local maps are all-gathered and compared with the checksums of the gathered rows; the verdicts
    are MIN-reduced.  (bench.py runs this once per multi-GPU run; tests/test_distributed_cpu.py on gloo.)"""
    mine = label_checksums(local_u8.reshape(1, -1))[0]
    sums = _all_gather_rows(mine, group)
    ok = torch.tensor([int(torch.equal(label_checksums(gathered_u8), sums))], device=local_u8.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return bool(ok.item())


def segment_frames(images, depths, network, first_indices, cluster_fn, rank=0, world_size=1, pad_to=None):
    """Shard [F,3,H,W] host frames over ranks, run network + clustering on the local shard with the
    pre-drawn first-seed indices, all-gather the label maps.  `cluster_fn(features, firsts)` returns
    integer labels [F_local, H*W] on the feature device."""
    F_total = images.shape[0]
    b, e = shard_range(F_total, world_size, rank)
    per = (F_total + world_size - 1) // world_size
    H, W = images.shape[2], images.shape[3]
    outs = []
    for f in range(b, e):
        feats = network(images[f:f + 1], None, depths[f:f + 1])
        outs.append(cluster_fn(feats, [first_indices[f]]).view(1, H, W))
    dev = outs[0].device if outs else torch.device("cpu")
    while len(outs) < per:                              # tail rank padding so that every rank sends `per` maps
        outs.append(torch.zeros((1, H, W), dtype=torch.int32, device=dev))
    local = torch.cat(outs, 0).to(torch.int32)
    return gather_labels(local, F_total)
