#!/usr/bin/env python
"""bench.py -- frames/s of the UnseenObjectClustering hot path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over --batch synthetic 640x480 RGB-D frames per GPU:
  backbone (two-branch ResNet34-8s, tcgen05 implicit-GEMM convs, fused head) -> stage-1 clustering
  (farthest point sampling, 10 tcgen05 mean-shift updates, seed labelling, pixel labels)
= BASELINE.json configs[1] ("640x480 RGB-D, 64-dim cosine embeddings, stage-1 mean-shift only").

  python bench.py --gpus N --steps K --warmup W            # this repo (N>1: launched by torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0):
  value                frames/s, inputs resident in HBM (pipelined: --depth steps of --batch frames in flight)
  e2e                  the same from RAW host frames (uint8 BGR + uint16 depth, pinned): H2D + device input preparation
                       (tools/test_images.py:105-135) + D2H of the float32 label maps inside the timed region
  e2e_fp32_inputs      the same from the reference API's fp32 sample tensors (7.4 MB / frame over PCIe)
  roofline             the mean-shift loop kernel on SURVEY 8(d)'s config-2 clustered field against the measured HBM peak
                       (B = 1 launch, and the per-frame figure of a 4-field launch)
  config2_clustered    stage times of the stage-1 clustering on that field;  stages_ms: the same on the backbone's own field
  config3 / config5    BASELINE configs 3 (two-stage through test_sample) and 5 (960x720x128, 30 updates: HBM-streamed)
  sustained            >= 3 s of the pipelined workload without a pause: frames/s, clock median, power
  torch_gpu_baseline   the reference's PyTorch path on this GPU (cuDNN / cuBLAS eager, oracle/uoc_torch_gpu.py) and the ratio
                       against it (north_star: >= 30x)
  cpu_baseline         the reference's CPU path (oracle port) on the host cores
  N > 1: frames sharded over ranks, uint8 label maps all-gathered over NCCL on a side stream, gathered result verified
  (config4: 8 frames per GPU; config5 on N GPUs: one 960x720x128 field per GPU).
Timing: CUDA events on the launching streams, max over ranks; L2 is flushed with a 256 MiB memset before EVERY step,
inside the timed region; `serial` is the one-frame-at-a-time latency (flush untimed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    # NCCL prints its version banner to STDOUT at the VERSION (and WARN) level; stdout is meant to be the one JSON line.
    # INFO / TRACE settings are left alone.
    os.environ.pop("NCCL_DEBUG")

H, W, D, M, ITERS, KAPPA = 480, 640, 64, 100, 10, 20.0
METRIC = "frames/sec 640x480 RGB-D seg"
UNIT = "frames/s"
WORKLOAD = "640x480 RGB-D, ResNet34-8s x2 (random init), 64-dim cosine embeddings, stage-1 mean-shift (100 seeds, 10 iters)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}




def config_dict(args, world):
    """The SAME dictionary in both arms (b200 / reference): the driver compares them."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": max(1, args.batch), "steps_in_flight": max(1, args.depth),
            "l2": "flushed before every step (256 MiB memset on the step's stream, inside the timed region)",
            "timing": "CUDA events on the launching streams (start -> last frame done), max over ranks",
            "multi_gpu": "frames sharded over ranks, one NCCL all-gather of uint8 label maps per step" if world > 1 else "single GPU"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def best_cpu_threads(budget_s=45.0):
    """The reference runs torch with its default thread count; on a many-core host that is far from its best.
    Probe a few counts on one frame each (bounded) and return the fastest -- the baseline gets every advantage."""
    ncpu = os.cpu_count() or 1
    cands = [c for c in (16, 32, 8) if 1 <= c <= ncpu]      # 64 / 128 threads were measured 5-20x slower (oversubscribed small ops)
    cands = list(dict.fromkeys(cands)) or [ncpu]
    best, best_t = cands[0], None
    t_start = time.perf_counter()
    for c in cands:
        if time.perf_counter() - t_start > budget_s:
            break
        _, _, per = cpu_reference_frames(1, 0, c)
        if best_t is None or per < best_t:
            best, best_t = c, per
    return best


def cpu_reference_frames(steps, warmup, threads=None):
    """The reference's CPU path through the oracle port (oracle/uoc_oracle.py: same torch CPU ops as
    the reference, pinned bit-identically against it): backbone forward + clustering_features on one
    640x480 frame per step.  Returns (frames/s, torch threads, seconds per frame)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import uoc_oracle as O
    from unseenobjectclustering_b200.networks import random_state_dict
    if threads is None:
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = O.OracleSegNet(random_state_dict(D, seed=0))
    img, xyz = O.synthetic_rgbd_frame(H, W, seed=0)
    np.random.seed(3)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        feats = net(img, None, xyz)
        O.clustering_features(feats, M)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    return 1.0 / per, torch.get_num_threads(), per


CPU_NOTE = ("oracle port = the reference's own torch CPU ops (pinned bit-identical to it); runs under torch.no_grad() whereas the "
            "reference keeps autograd on at inference (SURVEY 9.12): the baseline is, if anything, flattered")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 3)
    budget_s = 280.0
    t0 = time.perf_counter()
    threads = best_cpu_threads(20.0)
    fps1, cores, per = cpu_reference_frames(1, 1, threads)  # probe the per-frame cost
    room = int((budget_s - (time.perf_counter() - t0)) / max(per, 1e-3))
    if warmup + steps > room:                               # a slow host: bounded sample, said so in the line
        warmup = min(warmup, 3)
        steps = max(1, room - warmup)
    fps, cores, per = cpu_reference_frames(steps, warmup, threads)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world),
        "reference_arm": "the reference's CPU path (torch CPU ops of the oracle port) on the host cores, ONE frame per step",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": "port",
                         "sample": "%d timed full frames (backbone + stage-1 clustering) after %d warm-up; torch threads = best of a "
                                   "probe over {16, 32, 8}" % (steps, warmup), "note": CPU_NOTE},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def stage_split(lib, _lib, MS, dev, feats, xb, n, d, m, iters, firsts, flush, reps, batch=1, backbone=None):
    """Stage times (CUDA events, L2 flushed before every repetition, median) of the stage-1 clustering through the stage
    entry points of the C ABI: {fps, loop, label_assign} (+ backbone when `backbone` = (net, img, xyz))."""
    import ctypes
    import torch
    ws = MS._workspace(dev, lib.uoc_meanshift_workspace_bytes(batch, n, d, m))
    sel = torch.empty((batch, m), dtype=torch.int64, device=dev)
    Z = torch.empty((batch, m, d), dtype=torch.float32, device=dev)
    sl = torch.empty((batch, m), dtype=torch.int32, device=dev)
    nu = torch.empty((batch,), dtype=torch.int32, device=dev)
    lab = torch.empty((batch, n), dtype=torch.int32, device=dev)
    sp = _lib.stream_ptr(dev)
    keys = (["backbone"] if backbone else []) + ["fps", "loop", "label_assign"]
    acc = {k: [] for k in keys}
    for rep in range(reps + 1):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(keys) + 1)]
        k = 0
        if backbone:
            net, a, b = backbone
            torch.cuda._sleep(8_000_000)      # ~4 ms of GPU spin: the host enqueues the backbone's launches meanwhile (else host-bound)
            ev[0].record()
            feats, xb = net.forward_ex(a, None, b)
            k = 1
        ev[k].record()
        first = (ctypes.c_int64 * batch)(*[int(firsts[(rep * batch + j) % len(firsts)]) for j in range(batch)])
        sb, sd = feats.stride(0), feats.stride(1)
        _lib.check(lib.uoc_select_seeds(_lib.ptr(feats), sb, sd, _lib.ptr(xb), batch, n, d, m, ctypes.cast(first, ctypes.c_void_p),
                                        _lib.ptr(sel), _lib.ptr(Z), _lib.ptr(ws), ws.numel(), 0, sp), "select_seeds")
        ev[k + 1].record()
        _lib.check(lib.uoc_hill_climb(_lib.ptr(feats), sb, sd, _lib.ptr(xb), batch, n, d, m, KAPPA, iters, _lib.ptr(Z),
                                      _lib.ptr(ws), ws.numel(), 0, sp), "hill_climb")
        ev[k + 2].record()
        _lib.check(lib.uoc_label_seeds(_lib.ptr(Z), batch, m, d, 0.04, _lib.ptr(sl), _lib.ptr(nu), sp), "label_seeds")
        _lib.check(lib.uoc_assign_labels(_lib.ptr(feats), sb, sd, _lib.ptr(xb), batch, n, d, m, _lib.ptr(Z), _lib.ptr(sl), _lib.ptr(nu),
                                         _lib.ptr(lab), _lib.ptr(ws), ws.numel(), sp), "assign_labels")
        ev[k + 3].record()
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for key, (x, y) in zip(keys, zip(ev[:-1], ev[1:])):
            acc[key].append(x.elapsed_time(y))
    out = {key: sorted(v)[len(v) // 2] for key, v in acc.items()}
    out["clusters"] = int(lab[0].max().item()) + 1
    return out


def loop_roofline(ms, n, d, iters, peak, peak_src, traffic, kernel, fields=1):
    """Algorithmic bytes = the bf16 pixel-major copy streamed once per update: iters * n * d * 2 per field."""
    nbytes = iters * n * d * 2 * fields
    gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "peak_source": peak_src, "traffic": traffic, "bytes_per_launch": nbytes, "ms_per_launch": ms,
            "fp32_equivalent_GBps": 2.0 * gbs}


def torch_gpu_baseline(dev, frames, firsts, reps=3):
    """The reference's PyTorch path on this GPU (oracle/uoc_torch_gpu.py: cuDNN convolutions, cuBLAS mm, eager kernels,
    the reference's host round trips; device shim of SURVEY 0.10).  H2D of the fp32 sample and the float32 CPU label map
    are inside the timed region (wall clock around a synchronised frame: the path itself synchronises ~200 times)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import uoc_torch_gpu as G
    from unseenobjectclustering_b200.networks import random_state_dict
    sd = random_state_dict(D, seed=0)
    modes = {
        "default": dict(cudnn_tf32=True, matmul_tf32=False, grad=True,
                        note="torch defaults, autograd on: exactly how the reference runs (cuDNN TF32 convolutions, fp32 mm)"),
        "default_no_grad": dict(cudnn_tf32=True, matmul_tf32=False, grad=False, note="torch defaults under no_grad"),
        "strict_fp32": dict(cudnn_tf32=False, matmul_tf32=False, grad=False, note="TF32 off everywhere, no_grad"),
        "all_tf32": dict(cudnn_tf32=True, matmul_tf32=True, grad=False, note="TF32 allowed for convolutions and mm, no_grad"),
    }
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {}
    try:
        for name, mo in modes.items():
            torch.backends.cudnn.allow_tf32 = mo["cudnn_tf32"]
            torch.backends.cuda.matmul.allow_tf32 = mo["matmul_tf32"]
            net = G.TorchGpuSegNet(sd, dev, grad=mo["grad"])
            times = []
            for i in range(reps + 1):
                img, xyz = frames[i % len(frames)]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                G.frame(net, img, xyz, firsts[i], dev)
                torch.cuda.synchronize()
                if i > 0:
                    times.append(time.perf_counter() - t0)
            per = sorted(times)[len(times) // 2]
            out[name] = {"frames_per_s": 1.0 / per, "ms_per_frame": per * 1e3, "note": mo["note"]}
            del net
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def run_b200(args):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from unseenobjectclustering_b200 import _lib, networks, synthetic
    from unseenobjectclustering_b200 import mean_shift as MS
    from unseenobjectclustering_b200 import distributed as UD
    from unseenobjectclustering_b200 import test_dataset as TD
    from unseenobjectclustering_b200.pipeline import FramePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    steps, warmup = args.steps, max(args.warmup, 3)
    extras = world == 1 and not args.quick

    net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
    n = H * W
    nframes = 4                                            # distinct inputs, rotated
    frames = [synthetic.rgbd_frame(H, W, seed=100 * rank + i) for i in range(nframes)]
    dev_frames = [(a.to(dev), b.to(dev)) for a, b in frames]
    pin_frames = [(a.pin_memory(), b.pin_memory()) for a, b in frames]
    B = max(1, args.batch)                                 # frames per GPU per step (one slot launch of the pipeline)
    firsts = UD.draw_first_indices(4 * (warmup + 4 * steps + 64) * B + 4096, n, seed=3 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_pin = torch.empty((1, H, W), dtype=torch.float32).pin_memory()
    lab_f32 = torch.empty((1, H, W), dtype=torch.float32, device=dev)

    # ---- serial latency: one frame at a time through the public calls ------------------------------------------
    def step_device(i):
        a, b = dev_frames[i % nframes]
        feats, xb = net.forward_ex(a, None, b)
        labels, _ = MS.cluster_fields(feats, M, KAPPA, ITERS, [firsts[i]], x_bf16=xb)
        return labels

    def step_e2e(i):
        a, b = pin_frames[i % nframes]
        ad = a.to(dev, non_blocking=True)
        bd = b.to(dev, non_blocking=True)
        feats, xb = net.forward_ex(ad, None, bd)
        labels, _ = MS.cluster_fields(feats, M, KAPPA, ITERS, [firsts[i]], x_bf16=xb, labels_f32_out=lab_f32)
        out_pin.copy_(lab_f32, non_blocking=True)           # float32 CPU labels: the reference API
        return labels

    def timed(fn, count, base):
        evs = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for i in range(count):
            flush.zero_()                                   # L2 flush, outside the event pair
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(base + i)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(warmup):
        step_device(i)
        step_e2e(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    serial_steps = max(3, min(steps, 10))
    ms_dev = timed(step_device, serial_steps, warmup)
    ms_e2e = timed(step_e2e, serial_steps, warmup + serial_steps)
    # kernels of ONE step of B frames, counted by the library on a launch-by-launch pass (graph replays -- the public forward
    # and the pipeline -- are invisible to the counter; they replay exactly these launches)
    imgB = torch.cat([dev_frames[i % nframes][0] for i in range(B)], 0)
    xyzB = torch.cat([dev_frames[i % nframes][1] for i in range(B)], 0)
    l0 = lib.uoc_launch_count()
    fB, xB = net.forward_ex(imgB, None, xyzB, graph=False)
    MS.cluster_fields(fB, M, KAPPA, ITERS, [firsts[i] for i in range(B)], x_bf16=xB)
    torch.cuda.synchronize()
    launches_per_step = int(lib.uoc_launch_count() - l0)
    del imgB, xyzB, fB, xB

    # ---- pipelined throughput: --depth steps of B frames in flight on separate streams --------------------------
    pipe = FramePipeline(net, H, W, depth=max(1, args.depth), num_seeds=M, kappa=KAPPA, max_iters=ITERS, device=dev,
                         frames_per_slot=B)
    host_ms = [0.0]
    rng = np.random.RandomState(7 + rank)
    raw_frames = [(torch.from_numpy(rng.randint(0, 256, (H, W, 3)).astype(np.uint8)).pin_memory(),
                   torch.from_numpy(rng.randint(300, 1500, (H, W)).astype(np.int16)).pin_memory()) for _ in range(nframes)]
    camera = {"fx": 612.937, "fy": 613.173, "x_offset": 322.549, "y_offset": 248.158}   # data/demo/camera_params.json
    comm = torch.cuda.Stream(device=dev) if world > 1 else None        # the label gather runs beside the next step's kernels
    gathered_u8 = torch.empty((world, B * n), dtype=torch.uint8, device=dev) if world > 1 else None
    gather_done = {}

    def timed_pipe(mode, count, base):
        """mode: 'resident' | 'fp32' (pinned fp32 sample tensors) | 'raw' (pinned uint8 / uint16 frames)."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        for sl in pipe.slots:
            sl.stream.wait_event(start)
        if comm is not None:
            comm.wait_event(start)
        ends = []
        host_t0 = time.perf_counter()
        for i in range(count * B):                          # count steps of B frames
            sl = pipe.slots[pipe.next % len(pipe.slots)]
            if sl.fill == 0:
                if sl.busy:
                    pipe.collect_one()
                with torch.cuda.stream(sl.stream):
                    if id(sl) in gather_done:
                        sl.stream.wait_event(gather_done.pop(id(sl)))   # the previous gather has read this slot's labels
                    flush.zero_()                           # L2 flush before every step, INSIDE the timed region
            if mode == "raw":
                a, b = raw_frames[(base + i) % nframes]
                pipe.submit_raw(a, b, camera, firsts[base + i])
            else:
                a, b = (dev_frames if mode == "resident" else pin_frames)[(base + i) % nframes]
                pipe.submit(a, b, firsts[base + i], resident=(mode == "resident"))
            if sl.fill != 0:
                continue                                    # the slot's batch is not complete yet
            e = torch.cuda.Event(enable_timing=True)
            if world > 1:
                # uint8 label maps (written by the label pass itself) gathered on a side stream behind an event
                comm.wait_event(sl.done)
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(gathered_u8, sl.lab_u8.view(-1))
                    e.record(comm)
                gather_done[id(sl)] = e
            else:
                e.record(sl.stream)
            ends.append(e)
        host_ms[0] = (time.perf_counter() - host_t0) * 1e3 / max(count, 1)     # host time to enqueue one step
        pipe.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = max(start.elapsed_time(e) for e in ends)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(warmup):
        for mode in ("resident", "fp32", "raw"):
            timed_pipe(mode, 2, i)
    l0 = lib.uoc_launch_count()
    ms_pipe_dev = timed_pipe("resident", steps, warmup)
    host_enqueue_ms = host_ms[0]
    launches_pipe = int(lib.uoc_launch_count() - l0)
    ms_pipe_raw = timed_pipe("raw", steps, warmup + steps)
    ms_pipe_fp32 = timed_pipe("fp32", steps, warmup + 2 * steps)

    # ---- N > 1: the gathered labels are checked once per run; BASELINE configs 4 and 5 on N GPUs --------------
    multi = None
    if world > 1:
        multi = {}
        sl = pipe.slots[(pipe.next - 1) % len(pipe.slots)]          # the last step: its labels and the last gather
        from unseenobjectclustering_b200 import distributed as UD
        arrived = UD.verify_gathered_labels(sl.lab_u8.view(-1), gathered_u8)     # checksums all-gathered, verdict MIN-reduced
        same_as_f32 = torch.tensor([int(torch.equal(sl.lab_u8.view(B, H, W).to(torch.float32), sl.lab_f32))], device=dev)
        dist.all_reduce(same_as_f32, op=dist.ReduceOp.MIN)
        multi["gather_verified"] = bool(arrived and same_as_f32.item())
        multi["gather_dtype"] = "uint8 on the wire (%d bytes per frame), widened by the receiver" % n
        # config 4: 8 frames per GPU (8 / S steps of S), ONE gather of all 8*N label maps, inside the timed region
        per_gpu = 8
        S = B if per_gpu % B == 0 else 4
        pipe4 = pipe if S == B else FramePipeline(net, H, W, depth=max(2, args.depth), num_seeds=M, kappa=KAPPA, max_iters=ITERS,
                                                  device=dev, frames_per_slot=S)
        local8 = torch.empty((per_gpu, n), dtype=torch.uint8, device=dev)
        all8 = torch.empty((world, per_gpu * n), dtype=torch.uint8, device=dev)
        t4 = []
        for rep in range(4):
            dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for psl in pipe4.slots:
                psl.stream.wait_event(s)
            done = []
            for i in range(per_gpu):
                a, b = raw_frames[i % nframes]
                psl = pipe4.slots[pipe4.next % len(pipe4.slots)]
                pipe4.submit_raw(a, b, camera, firsts[1000 + i])
                if psl.fill == 0:
                    k = (i // S) * S
                    with torch.cuda.stream(psl.stream):
                        local8[k:k + S].copy_(psl.lab_u8.view(S, n), non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(psl.stream)
                    done.append(ev)
            for ev in done:
                torch.cuda.current_stream().wait_event(ev)
            dist.all_gather_into_tensor(all8, local8.view(-1))
            e.record()
            pipe4.drain()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                t4.append(float(t.item()))
        ms4 = sorted(t4)[len(t4) // 2]
        multi["config4"] = {"workload": "%d frames, 8 per GPU, raw host frames in, ONE NCCL all-gather of %d uint8 label maps" % (per_gpu * world, per_gpu * world),
                            "ms": ms4, "frames_per_s": per_gpu * world / (ms4 * 1e-3)}
        # config 5 on N GPUs: one 960x720x128 field per GPU (independent frames, weak scaling), whole stage-1 clustering
        H5, W5, D5, T5 = 720, 960, 128, 30
        f5, _ = synthetic.clustered_features(H5, W5, D5, 12, 0.05, seed=rank)
        f5 = f5.to(dev)
        x5 = MS.pack_bf16(f5)
        t5 = []
        for rep in range(4):
            flush.zero_()
            dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            MS.cluster_fields(f5, M, KAPPA, T5, [firsts[rep]], x_bf16=x5)
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                t5.append(float(t.item()))
        ms5 = sorted(t5)[len(t5) // 2]
        multi["config5"] = {"workload": "960x720, 128-dim embeddings, 30 mean-shift updates: one field per GPU, stage-1 clustering",
                            "ms_per_field": ms5, "fields_per_s": world / (ms5 * 1e-3)}
        del f5, x5

    # ---- stage split on the backbone's own field (same kernels through the stage entry points), rank-local ----
    reps = max(3, min(steps, 10))
    a0, b0 = dev_frames[0]
    st_bb = stage_split(lib, _lib, MS, dev, None, None, n, D, M, ITERS, firsts, flush, reps, 1, backbone=(net, a0, b0))
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    traffic = traffic4 = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = tj["meanshift_tc_persistent_kernel<64>"]["dram_bytes_per_launch"]
        traffic4 = tj["meanshift_tc_persistent_kernel<64>, 4 fields per launch"]["dram_bytes_per_launch"]
    except Exception:
        pass
    kname = "meanshift_tc_persistent_kernel<64>, one launch = all %d mean-shift updates" % ITERS
    line = {
        "metric": METRIC, "value": world * B * steps / (ms_pipe_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_pipe_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": config_dict(args, world),
        "clocks": clocks,
        # the user's call: raw frames as cv2.imread returns them (tools/test_images.py:105-135 arithmetic on the device)
        "e2e": {"value": world * B * steps / (ms_pipe_raw * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * H * W * 5,
                "d2h_bytes_per_step": B * H * W * 4, "ms_per_step": ms_pipe_raw / steps,
                "inputs": "raw uint8 BGR + uint16 depth frames in pinned host memory; float32 CPU label maps out"},
        "e2e_fp32_inputs": {"value": world * B * steps / (ms_pipe_fp32 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 2 * 3 * H * W * 4,
                            "d2h_bytes_per_step": B * H * W * 4, "ms_per_step": ms_pipe_fp32 / steps,
                            "inputs": "the reference API's fp32 sample tensors (image_color, depth XYZ) in pinned host memory"},
        "serial": {"value": world * serial_steps / (ms_dev * 1e-3), "ms_per_step": ms_dev / serial_steps,
                   "e2e_value": world * serial_steps / (ms_e2e * 1e-3), "e2e_ms_per_step": ms_e2e / serial_steps,
                   "note": "one frame at a time, L2 flushed (untimed) between frames"},
        # the pipelined region replays CUDA graphs (not visible to the library's launch counter): the same kernels as the
        # launch-by-launch pass of one step above, times the timed steps
        "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step,
        "gpu_launches_eager_in_pipeline": launches_pipe,
        "host_loop_ms_per_step": round(host_enqueue_ms, 3), "host_cores": host_cores(),
        "stages_ms": {k: round(v, 4) for k, v in st_bb.items() if k != "clusters"},
    }
    if multi is not None:
        line["multi_gpu"] = multi

    # ---- config 2 on SURVEY 8(d)'s clustered field: the roofline of the loop kernel --------------------------
    fc, _ = synthetic.clustered_features(H, W, D, 6, 0.05, seed=0)
    fc = fc.to(dev)
    xc = MS.pack_bf16(fc)
    st_c = stage_split(lib, _lib, MS, dev, fc, xc, n, D, M, ITERS, firsts, flush, reps, 1)
    line["config2_clustered"] = {"workload": "SURVEY 8(d) config 2: 640x480x64 unit field, 6 objects + background, noise 0.05",
                                 "stages_ms": {k: round(v, 4) for k, v in st_c.items() if k != "clusters"},
                                 "clusters": st_c["clusters"]}
    roof = loop_roofline(st_c["loop"], n, D, ITERS, peak, peak_src, traffic, kname)
    roof["field"] = "config-2 clustered field (L2-resident after the first update: DRAM traffic << algorithmic bytes)"
    roof["backbone_field"] = loop_roofline(st_bb["loop"], n, D, ITERS, peak, peak_src, None, kname)
    if extras:
        f4 = torch.cat([synthetic.clustered_features(H, W, D, 6, 0.05, seed=s)[0] for s in range(4)], 0).to(dev)
        x4 = MS.pack_bf16(f4)
        st_4 = stage_split(lib, _lib, MS, dev, f4, x4, n, D, M, ITERS, firsts, flush, max(3, reps // 2), 4)
        roof["batched4"] = loop_roofline(st_4["loop"], n, D, ITERS, peak, peak_src, traffic4, kname + ", 4 fields per launch", fields=4)
        roof["batched4"]["ms_per_frame"] = st_4["loop"] / 4
        roof["batched4"]["field"] = "four bf16 fields = 157 MB > L2: streamed from HBM in every update (DRAM traffic ~ algorithmic bytes)"
        line["config2_clustered"]["stages_ms_per_frame_4_fields"] = {k: round(v / 4, 4) for k, v in st_4.items() if k != "clusters"}
        del f4, x4
        if B not in (1, 4):
            # the launch shape of the timed region itself: B fields per persistent launch (frames_per_gpu_per_step)
            fB = torch.cat([synthetic.clustered_features(H, W, D, 6, 0.05, seed=s)[0] for s in range(B)], 0).to(dev)
            xB = MS.pack_bf16(fB)
            st_B = stage_split(lib, _lib, MS, dev, fB, xB, n, D, M, ITERS, firsts, flush, max(3, reps // 2), B)
            roof["as_in_timed_region"] = loop_roofline(st_B["loop"], n, D, ITERS, peak, peak_src, None,
                                                       kname + ", %d fields per launch" % B, fields=B)
            roof["as_in_timed_region"]["ms_per_frame"] = st_B["loop"] / B
            line["config2_clustered"]["stages_ms_per_frame_%d_fields" % B] = {k: round(v / B, 4) for k, v in st_B.items() if k != "clusters"}
            del fB, xB
    roof["note"] = "algorithmic bytes = n*d*2 per mean-shift update (the bf16 copy actually streamed); fp32-equivalent (n*d*4) is 2x"
    line["roofline"] = roof
    del fc, xc

    if extras:
        # ---- config 3: two-stage through test_sample (as tools/bench_configs.py: real forward passes, synthetic fields) ----
        net_crop = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=1)).to(dev)
        feats1 = synthetic.clustered_features(H, W, D, 6, 0.05, seed=0)[0].to(dev)
        g = torch.Generator().manual_seed(5)
        crop_centres = torch.nn.functional.normalize(torch.randn(2, D, generator=g), dim=1).to(dev)
        crop_noise = (0.05 * torch.randn(8, D, 224, 224, generator=g)).to(dev)

        def stage1(img, label, depth):
            net(img, label, depth)                      # the real forward pass (timed); its collapsed field is not used
            return feats1

        def stage2(img, label, depth):
            net_crop(img, label, depth)                 # the real crop forward pass (timed)
            ids = (label > 0).long()                    # crop fields that follow the stage-1 mask crops: object / rest
            f = crop_centres[ids].permute(0, 3, 1, 2) + crop_noise[:img.shape[0]]
            return torch.nn.functional.normalize(f, dim=1).contiguous()

        sample = {"image_color": pin_frames[0][0], "depth": pin_frames[0][1]}
        t3, objects = [], 0
        for rep in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out_label, refined = TD.test_sample(sample, stage1, stage2, [1000 + rep], [77 + k for k in range(8)])
            b.record()
            torch.cuda.synchronize()
            if rep >= 2:
                t3.append(a.elapsed_time(b))
            objects = int(refined.max().item()) if refined is not None else 0
        ms3 = sorted(t3)[len(t3) // 2]
        line["config3"] = {"workload": "640x480 RGB-D two-stage (6 crops of 224x224), one frame at a time through test_sample "
                                       "(fp32 host sample in, float32 CPU label maps out)", "ms_per_frame": ms3,
                           "frames_per_s": 1000.0 / ms3, "objects_after_refinement": objects}
        del net_crop, feats1, crop_noise

        # ---- config 5: 960x720x128, 30 updates: the field does not fit in L2 -> streamed from HBM in every update ----
        H5, W5, D5, T5 = 720, 960, 128, 30
        n5 = H5 * W5
        f5 = synthetic.clustered_features(H5, W5, D5, 12, 0.05, seed=0)[0].to(dev)
        x5 = MS.pack_bf16(f5)
        st_5 = stage_split(lib, _lib, MS, dev, f5, x5, n5, D5, M, T5, [n5 // 3, n5 // 5, n5 // 7], flush, 4, 1)
        roof5 = loop_roofline(st_5["loop"], n5, D5, T5, peak, peak_src, None,
                              "meanshift_tc_persistent_kernel<128>, one launch = all 30 mean-shift updates")
        roof5["note"] = "bf16 field 177 MB > L2: streamed from HBM in every update"
        line["config5"] = {"workload": "960x720, 128-dim embeddings, 100 seeds, 30 mean-shift updates (one field per GPU)",
                           "stages_ms": {k: round(v, 4) for k, v in st_5.items() if k != "clusters"}, "clusters": st_5["clusters"],
                           "roofline": roof5}
        del f5, x5
        torch.cuda.empty_cache()

        # ---- sustained: >= 3 s of the pipelined workload (resident inputs) without a pause ----
        per_step_ms = ms_pipe_dev / steps
        sus_steps = max(steps, int(3300.0 / max(per_step_ms, 0.05)))
        sam2 = ClockSampler(local)
        sam2.start()
        ms_sus = timed_pipe("resident", sus_steps, 0)
        ck = sam2.stop()
        line["sustained"] = {"value": B * sus_steps / (ms_sus * 1e-3), "unit": UNIT, "seconds": ms_sus * 1e-3, "steps": sus_steps,
                             "ms_per_step": ms_sus / sus_steps, "clocks": ck}

        # ---- the reference's PyTorch path on this GPU, and the CPU path on the host cores ----
        try:
            tg = torch_gpu_baseline(dev, [(a, b) for a, b in pin_frames], firsts)
            best = max(v["frames_per_s"] for v in tg.values())
            as_run = tg["default"]["frames_per_s"]
            line["torch_gpu_baseline"] = {
                "kind": "port of the reference's GPU path (cuDNN / cuBLAS / eager torch ops, reference host round trips; device "
                        "shim of SURVEY 0.10)", "unit": UNIT, "modes": tg,
                "e2e_ratio_vs_reference_as_run": line["e2e_fp32_inputs"]["value"] / as_run,
                "e2e_ratio_vs_fastest_mode": line["e2e_fp32_inputs"]["value"] / best,
                "serial_e2e_ratio_vs_reference_as_run": line["serial"]["e2e_value"] / as_run,
                "target": "north_star: >= 30x the reference's single-GPU PyTorch frames/sec on the same box",
                "note": "both sides: fp32 sample tensors from host memory in, float32 CPU label map out"}
        except Exception as e:                       # the baseline must never take the bench line down
            line["torch_gpu_baseline"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        threads = best_cpu_threads(12.0)
        fps, cores, per = cpu_reference_frames(3, 1, threads)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": "port",
                                "sample": "3 timed full frames (oracle backbone + stage-1 clustering) after 1 warm-up, best torch "
                                          "thread count of a bounded probe, %.1f s in total" % (time.perf_counter() - t0),
                                "note": CPU_NOTE}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the extra legs (configs 3 / 5, sustained, torch GPU baseline)")
    ap.add_argument("--depth", type=int, default=2, help="steps in flight per GPU (1 = strictly serial)")
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step: they go through every kernel together")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
