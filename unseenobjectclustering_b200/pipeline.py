"""Frame pipeline: several frames in flight on separate CUDA streams.

One frame is a chain of dependent kernels (backbone -> sampling -> mean-shift loop -> labels) with
latency-bound stretches -- the farthest point sampling spends most of each pass waiting for an
inter-CTA exchange, tile-quantised convolutions leave SMs idle -- so a second frame on another
stream fills the gaps.  Each slot owns its stream, pinned host staging buffers, device inputs and
(through the per-stream workspaces of networks.py / mean_shift.py) its scratch memory.  Sampling
kernels are cooperative launches; they are chained across slots with events so that two of them never
compete for residency.
"""
import numpy as np
import torch

from . import mean_shift as _ms


class _Slot(object):
    def __init__(self, dev, H, W):
        self.stream = torch.cuda.Stream(device=dev)
        self.img_pin = torch.empty((1, 3, H, W), dtype=torch.float32).pin_memory()
        self.xyz_pin = torch.empty((1, 3, H, W), dtype=torch.float32).pin_memory()
        self.img_dev = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
        self.xyz_dev = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
        self.out_pin = torch.empty((1, H, W), dtype=torch.float32).pin_memory()
        self.done = torch.cuda.Event()
        self.labels = None
        self.busy = False


class FramePipeline(object):
    """submit() frames, collect float32 CPU label maps (the reference's out_label) in order."""

    def __init__(self, network, H=480, W=640, depth=2, num_seeds=100, kappa=20.0, max_iters=10, device=None):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.network = network
        self.H, self.W = H, W
        self.num_seeds, self.kappa, self.max_iters = num_seeds, kappa, max_iters
        self.slots = [_Slot(self.dev, H, W) for _ in range(depth)]
        self.next = 0
        self.coop_tail = None          # event after the last cooperative sampling kernel
        self.pending = []

    def _run(self, slot, img_dev, xyz_dev, first_index):
        feats = self.network(img_dev, None, xyz_dev)
        if self.coop_tail is not None:
            slot.stream.wait_event(self.coop_tail)          # sampling kernels never overlap each other
        def sampled():
            ev = torch.cuda.Event()
            ev.record(slot.stream)          # right after this frame's sampling kernel: the next frame's may start
            self.coop_tail = ev

        labels, _ = _ms.cluster_fields(feats, self.num_seeds, self.kappa, self.max_iters, [int(first_index)],
                                       on_sampling_done=sampled)
        return labels

    def submit(self, image, depth, first_index=None, resident=False):
        """image / depth: [1,3,H,W] float32 CPU tensors (or device tensors with resident=True)."""
        slot = self.slots[self.next % len(self.slots)]
        self.next += 1
        if slot.busy:
            self.collect_one()
        if first_index is None:
            first_index = np.random.randint(0, self.H * self.W)     # lib/utils/mean_shift.py:155, drawn in frame order
        with torch.cuda.stream(slot.stream):
            if resident:
                img_dev, xyz_dev = image, depth
            else:
                src_i, src_d = image, depth
                if not image.is_pinned():                  # stage through the slot's pinned buffers
                    slot.img_pin.copy_(image)
                    src_i = slot.img_pin
                if not depth.is_pinned():
                    slot.xyz_pin.copy_(depth)
                    src_d = slot.xyz_pin
                slot.img_dev.copy_(src_i, non_blocking=True)
                slot.xyz_dev.copy_(src_d, non_blocking=True)
                img_dev, xyz_dev = slot.img_dev, slot.xyz_dev
            slot.labels = self._run(slot, img_dev, xyz_dev, first_index)
            if not resident:
                slot.out_pin.copy_(slot.labels.view(1, self.H, self.W).to(torch.float32), non_blocking=True)
            slot.done.record(slot.stream)
        slot.busy = True
        self.pending.append(slot)
        return slot

    def collect_one(self):
        slot = self.pending.pop(0)
        slot.done.synchronize()
        slot.busy = False
        return slot.out_pin, slot.labels

    def drain(self):
        out = []
        while self.pending:
            out.append(self.collect_one())
        return out
