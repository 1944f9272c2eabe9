"""Frame pipeline: several frames in flight on separate CUDA streams, replayed from CUDA graphs.

One frame is a chain of ~65 dependent kernels (backbone -> sampling -> mean-shift loop -> labels) with
latency-bound stretches -- the farthest point sampling spends most of each pass waiting for an
inter-CTA exchange, tile-quantised convolutions leave SMs idle -- so a second frame on another
stream fills the gaps.  Each slot owns its stream, device inputs, outputs and scratch memory.

Enqueueing a frame eagerly costs ~1.7 ms of host time (ctypes calls, ~150 tensor-map encodes,
~70 launches), which caps the throughput; so after a warm-up every slot captures TWO CUDA graphs:
  A: backbone forward   (inputs: the slot's device frame buffers; outputs: features + bf16 copy)
  B: seed labelling + pixel labels + D2H of the label map
The sampling kernel and the persistent mean-shift loop kernel between them stay eager cooperative
launches (2 launches): they are chained across slots with an event so that cooperative kernels of
different frames never compete for residency; the first-seed index is a per-frame host value.

frames_per_slot = B > 1: a slot collects B frames and runs them through every kernel together (backbone at
batch B, ONE persistent mean-shift launch and one label pass for the B fields; the sampling kernel keeps a whole
field resident on chip, so the B fields take turns there).  The tile-quantised convolutions and the exchange
steps of the persistent loop are shared between the frames: more frames per second, at B times the latency.
Results per frame are identical to B = 1 (batch items are independent in every kernel).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import mean_shift as _ms


class _Slot(object):
    def __init__(self, dev, H, W, B=1):
        self.stream = torch.cuda.Stream(device=dev)
        self.img_dev = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        self.xyz_dev = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        self.img_pin = None
        self.xyz_pin = None
        self.raw_im = None
        self.raw_dp = None
        self.out_pin = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
        self.fill = 0                # frames copied into the slot so far (frames_per_slot > 1)
        self.firsts = []
        self.done = torch.cuda.Event()
        self.err_pin = torch.zeros((1,), dtype=torch.int32).pin_memory()   # copy of the device error word, read after `done`
        self.labels = None
        self.lab_f32 = None
        self.lab_u8 = None
        self.busy = False
        self.graph_a = None
        self.graph_b = None
        self.runs = 0


class FramePipeline(object):
    """submit() frames, collect float32 CPU label maps (the reference's out_label) in order."""

    def __init__(self, network, H=480, W=640, depth=2, num_seeds=100, kappa=20.0, max_iters=10, device=None,
                 use_graphs=True, epsilon=None, frames_per_slot=1):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.network = network
        self.H, self.W = H, W
        self.B = max(1, int(frames_per_slot))
        self.num_seeds, self.kappa, self.max_iters = num_seeds, float(kappa), int(max_iters)
        self.eps = _ms._epsilon(epsilon)
        self.slots = [_Slot(self.dev, H, W, self.B) for _ in range(depth)]
        self.next = 0
        self.coop_tail = None          # event after the last cooperative sampling kernel
        self.pending = []
        self.collected = []            # results of slots that had to be recycled before the caller collected them
        self.use_graphs = bool(use_graphs) and isinstance(network, torch.nn.Module)
        self.graph_error = None

    # -- eager path (also the warm-up of the graph path) ----------------------------------------------
    def _forward(self, slot):
        """(features, bf16 copy or None): the network's explicit two-output form when it has one (SEGNET_B200)."""
        net = getattr(self.network, "module", self.network)
        if hasattr(net, "forward_ex"):
            return net.forward_ex(slot.img_dev, None, slot.xyz_dev, graph=False)     # the pipeline captures its own graphs
        feats = self.network(slot.img_dev, None, slot.xyz_dev)
        return feats, _ms._lookup_bf16(feats)

    def _run_eager(self, slot, firsts):
        feats, xb = self._forward(slot)
        if self.coop_tail is not None:
            slot.stream.wait_event(self.coop_tail)          # sampling kernels never overlap each other

        def sampled():
            ev = torch.cuda.Event()
            ev.record(slot.stream)          # right after this frame's sampling kernel: the next frame's may start
            self.coop_tail = ev

        if getattr(slot, "lab_f32", None) is None:
            slot.lab_f32 = torch.empty((self.B, self.H, self.W), dtype=torch.float32, device=self.dev)
            slot.lab_u8 = torch.empty((self.B, self.H * self.W), dtype=torch.uint8, device=self.dev)
        labels, _ = _ms.cluster_fields(feats, self.num_seeds, self.kappa, self.max_iters, [int(v) for v in firsts],
                                       epsilon=self.eps, on_sampling_done=sampled, x_bf16=xb, labels_f32_out=slot.lab_f32,
                                       labels_u8_out=slot.lab_u8)
        return labels

    # -- graph path -------------------------------------------------------------------------------------
    def _capture(self, slot):
        lib = _lib.load()
        dev, n, m, B = self.dev, self.H * self.W, self.num_seeds, self.B
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, stream=slot.stream):
            slot.feats, slot.xb = self._forward(slot)
        C = slot.feats.shape[1]
        slot.C = C
        slot.sel = torch.empty((B, m), dtype=torch.int64, device=dev)
        slot.Z = torch.empty((B, m, C), dtype=torch.float32, device=dev)
        slot.sl = torch.empty((B, m), dtype=torch.int32, device=dev)
        slot.nu = torch.empty((B,), dtype=torch.int32, device=dev)
        slot.lab = torch.empty((B, n), dtype=torch.int32, device=dev)
        slot.lab_f32 = torch.empty((B, self.H, self.W), dtype=torch.float32, device=dev)
        slot.lab_u8 = torch.empty((B, n), dtype=torch.uint8, device=dev)       # wire type of the multi-GPU label gather
        nbytes = lib.uoc_meanshift_workspace_bytes(B, n, C, m)
        slot.ws_fps = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=dev)
        slot.ws_b = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=dev)
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, stream=slot.stream):
            sp = _lib.stream_ptr(dev)
            f = slot.feats
            _lib.check(lib.uoc_label_seeds(_lib.ptr(slot.Z), B, m, C, self.eps, _lib.ptr(slot.sl), _lib.ptr(slot.nu), sp),
                       "uoc_label_seeds")
            _lib.check(lib.uoc_assign_labels_typed(_lib.ptr(f), f.stride(0), f.stride(1), _lib.ptr(slot.xb), B, n, C, m,
                                                   _lib.ptr(slot.Z), _lib.ptr(slot.sl), _lib.ptr(slot.nu), _lib.ptr(slot.lab),
                                                   _lib.ptr(slot.lab_f32), _lib.ptr(slot.lab_u8), _lib.ptr(slot.ws_b),
                                                   slot.ws_b.numel(), 0, sp), "uoc_assign_labels")
            slot.out_pin.copy_(slot.lab_f32, non_blocking=True)
            _lib.check(lib.uoc_peek_device_error_async(ctypes.c_void_p(slot.err_pin.data_ptr()), sp), "uoc_peek_device_error_async")
        slot.graph_a, slot.graph_b = ga, gb

    def _run_graph(self, slot, firsts):
        lib = _lib.load()
        n, m, C, B = self.H * self.W, self.num_seeds, slot.C, self.B
        slot.graph_a.replay()
        if self.coop_tail is not None:
            slot.stream.wait_event(self.coop_tail)
        first = (ctypes.c_int64 * B)(*[int(v) for v in firsts])
        f = slot.feats
        _lib.check(lib.uoc_select_seeds(_lib.ptr(f), f.stride(0), f.stride(1), _lib.ptr(slot.xb), B, n, C, m,
                                        ctypes.cast(first, ctypes.c_void_p),
                                        _lib.ptr(slot.sel), _lib.ptr(slot.Z), _lib.ptr(slot.ws_fps), slot.ws_fps.numel(), 0,
                                        _lib.stream_ptr(self.dev)), "uoc_select_seeds")
        # the mean-shift loop is ONE persistent cooperative kernel (all updates): eager as well
        _lib.check(lib.uoc_hill_climb(_lib.ptr(f), f.stride(0), f.stride(1), _lib.ptr(slot.xb), B, n, C, m, self.kappa,
                                      self.max_iters, _lib.ptr(slot.Z), _lib.ptr(slot.ws_b), slot.ws_b.numel(), 0,
                                      _lib.stream_ptr(self.dev)), "uoc_hill_climb")
        ev = torch.cuda.Event()
        ev.record(slot.stream)
        self.coop_tail = ev
        slot.graph_b.replay()
        return slot.lab

    # -- public -----------------------------------------------------------------------------------------
    def submit_raw(self, im_bgr, depth_raw, camera_params, first_index=None):
        """Raw frame: im_bgr [H,W,3] uint8 and depth_raw [H,W] uint16 (int16 storage) CPU tensors -- pinned for an
        asynchronous upload -- as cv2.imread returns them; the network inputs are built on the device by
        input_prep.prepare_inputs (tools/test_images.py:105-135 arithmetic, bit-identical): 1.5 MB instead of 7.4 MB
        over PCIe per frame."""
        return self.submit(None, None, first_index=first_index, raw=(im_bgr, depth_raw, camera_params))

    def submit(self, image, depth, first_index=None, resident=False, raw=None):
        """image / depth: [1,3,H,W] float32 tensors: pinned or pageable CPU tensors, or device tensors
        (resident=True).  Returns the slot; collect results with collect_one() / drain() in submission order."""
        slot = self.slots[self.next % len(self.slots)]
        if slot.busy:
            # the caller submitted more steps than there are slots without collecting: keep the oldest result (a copy --
            # the slot's pinned buffer is about to be reused) and hand it out from the next collect_one() / drain()
            out_pin, labels = self._collect_slot()
            self.collected.append((out_pin.clone(), labels.clone()))
        if first_index is None:
            first_index = np.random.randint(0, self.H * self.W)     # lib/utils/mean_shift.py:155, drawn in frame order
        k = slot.fill                                               # position of this frame inside the slot's batch
        if k == 0:
            slot.firsts = []
        slot.firsts.append(int(first_index))
        with torch.cuda.stream(slot.stream):
            if raw is not None:
                from . import input_prep
                im_bgr, depth_raw, cam = raw
                if slot.raw_im is None:
                    slot.raw_im = torch.empty((1, self.H, self.W, 3), dtype=torch.uint8, device=self.dev)
                    slot.raw_dp = torch.empty((1, self.H, self.W), dtype=torch.int16, device=self.dev)
                slot.raw_im.copy_(im_bgr.view(1, self.H, self.W, 3), non_blocking=True)
                slot.raw_dp.copy_(depth_raw.view(1, self.H, self.W), non_blocking=True)
                input_prep.prepare_inputs(slot.raw_im, slot.raw_dp, cam, device=self.dev,
                                          out=(slot.img_dev[k:k + 1], slot.xyz_dev[k:k + 1]))
            elif resident:
                slot.img_dev[k:k + 1].copy_(image, non_blocking=True)
                slot.xyz_dev[k:k + 1].copy_(depth, non_blocking=True)
            else:
                src_i, src_d = image, depth
                if not image.is_pinned():                  # stage through the slot's pinned buffers
                    if slot.img_pin is None:
                        slot.img_pin = torch.empty_like(slot.img_dev, device="cpu").pin_memory()
                        slot.xyz_pin = torch.empty_like(slot.xyz_dev, device="cpu").pin_memory()
                    slot.img_pin[k:k + 1].copy_(image)
                    slot.xyz_pin[k:k + 1].copy_(depth)
                    src_i, src_d = slot.img_pin[k:k + 1], slot.xyz_pin[k:k + 1]
                slot.img_dev[k:k + 1].copy_(src_i, non_blocking=True)
                slot.xyz_dev[k:k + 1].copy_(src_d, non_blocking=True)
            slot.fill += 1
            if slot.fill < self.B:
                return slot                     # the slot's batch is not complete yet: nothing is launched
            slot.fill = 0
            self.next += 1
            if self.use_graphs and slot.graph_a is None and slot.runs >= 1 and self.graph_error is None:
                try:
                    slot.stream.synchronize()
                    self._capture(slot)
                except Exception as e:           # capture is an optimisation of the host side only
                    self.graph_error = repr(e)
                    slot.graph_a = slot.graph_b = None
            if slot.graph_a is not None:
                slot.labels = self._run_graph(slot, slot.firsts)
            else:
                slot.labels = self._run_eager(slot, slot.firsts)
                slot.out_pin.copy_(slot.lab_f32, non_blocking=True)
                self._peek_error(slot)
            slot.runs += 1
            slot.done.record(slot.stream)
        slot.busy = True
        self.pending.append(slot)
        return slot

    def _peek_error(self, slot):
        _lib.check(_lib.load().uoc_peek_device_error_async(ctypes.c_void_p(slot.err_pin.data_ptr()), _lib.stream_ptr(self.dev)),
                   "uoc_peek_device_error_async")

    def _collect_slot(self):
        slot = self.pending.pop(0)
        slot.done.synchronize()
        slot.busy = False
        word = int(slot.err_pin[0])
        if word != 0:                  # a kernel of this step timed out / rejected its configuration: the labels are garbage
            with torch.cuda.device(self.dev):
                _lib.load().uoc_check_device_error(_lib.stream_ptr(self.dev))      # clears the word
            raise _lib.UocError("device-side pipeline error word 0x%x in a FramePipeline step (1=mbarrier time-out, "
                                "2=grid-barrier time-out, 4=bad config)" % word)
        return slot.out_pin, slot.labels

    def collect_one(self):
        """Oldest uncollected step: (float32 CPU label maps [B,H,W] -- the slot's pinned buffer, valid until the slot is
        reused `depth` steps later: copy it if it must live longer --, int32 device labels [B, H*W])."""
        if self.collected:
            return self.collected.pop(0)
        return self._collect_slot()

    def flush(self):
        """frames_per_slot > 1: run a partially filled slot now (the missing positions repeat the last frame and their
        results are to be ignored); returns the number of valid frames in it, 0 if nothing was waiting."""
        slot = self.slots[self.next % len(self.slots)]
        k = slot.fill
        if k == 0:
            return 0
        with torch.cuda.stream(slot.stream):
            for j in range(k, self.B):
                slot.img_dev[j:j + 1].copy_(slot.img_dev[k - 1:k], non_blocking=True)
                slot.xyz_dev[j:j + 1].copy_(slot.xyz_dev[k - 1:k], non_blocking=True)
                slot.firsts.append(slot.firsts[k - 1])
            slot.fill = 0
            self.next += 1
            slot.labels = (self._run_graph if slot.graph_a is not None else self._run_eager)(slot, slot.firsts)
            if slot.graph_a is None:
                slot.out_pin.copy_(slot.lab_f32, non_blocking=True)
                self._peek_error(slot)
            slot.runs += 1
            slot.done.record(slot.stream)
        slot.busy = True
        self.pending.append(slot)
        return k

    def drain(self):
        self.flush()
        out = []
        while self.collected or self.pending:
            out.append(self.collect_one())
        return out
