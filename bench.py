#!/usr/bin/env python
"""bench.py -- frames/s of the UnseenObjectClustering hot path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one synthetic 640x480 RGB-D frame per GPU:
  backbone (two-branch ResNet34-8s, tcgen05 implicit-GEMM convs, fused head) -> stage-1 clustering
  (farthest point sampling, 10 tcgen05 mean-shift updates, seed labelling, pixel labels)
= BASELINE.json configs[1] ("640x480 RGB-D, 64-dim cosine embeddings, stage-1 mean-shift only").

  python bench.py --gpus N --steps K --warmup W            # this repo (N>1: launched by torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same
through the public API with pinned HOST buffers (H2D of the frame + D2H of the labels inside the
timed region); `roofline` = the mean-shift loop kernel (the kernel BASELINE.json's metric names)
against the measured HBM peak; `cpu_baseline` = the oracle port on the host cores (N=1 only).
Timing: CUDA events on the launching streams, max over ranks.  A step = --batch frames per GPU that go
through every kernel together (pipeline.py, frames_per_slot); --depth steps are in flight on separate CUDA
streams; L2 is flushed with a 256 MiB memset before EVERY step, inside the timed region; `serial` in the
JSON line is the one-frame-at-a-time latency (flush untimed).
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
    640x480 frame per step.  Returns (frames/s, cores, seconds per frame)."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, min(args.warmup, 2)
    budget_s = 240.0
    t0 = time.perf_counter()
    threads = best_cpu_threads(20.0)
    fps1, cores, per = cpu_reference_frames(1, 1, threads)  # probe the per-frame cost
    steps_eff = max(1, min(steps, int((budget_s - (time.perf_counter() - t0)) / max(per, 1e-3)) - warmup))
    fps, cores, per = cpu_reference_frames(steps_eff, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_eff,
        "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path: torch CPU ops of the oracle port on the host cores"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d timed full frames (backbone + stage-1 clustering) after %d warm-up; torch threads = best of a "
                                   "probe over {all, half, 32, 16, 8} host cores" % (steps_eff, warmup)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from unseenobjectclustering_b200 import _lib, networks, synthetic
    from unseenobjectclustering_b200 import mean_shift as MS
    from unseenobjectclustering_b200 import distributed as UD

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

    net = networks.seg_resnet34_8s_embedding(2, D, networks.random_state_dict(D, seed=0)).to(dev)
    n = H * W
    nframes = 4                                            # distinct inputs, rotated
    frames = [synthetic.rgbd_frame(H, W, seed=100 * rank + i) for i in range(nframes)]
    dev_frames = [(a.to(dev), b.to(dev)) for a, b in frames]
    pin_frames = [(a.pin_memory(), b.pin_memory()) for a, b in frames]
    firsts = UD.draw_first_indices(4 * (warmup + 4 * steps + 16) * max(1, args.batch), n, seed=3 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = max(1, args.batch)                                 # frames per GPU per step (one slot launch of the pipeline)
    gathered = torch.empty((world, n), dtype=torch.int32, device=dev) if world > 1 else None
    gathered_b = torch.empty((world, B * n), dtype=torch.int32, device=dev) if world > 1 else None
    out_pin = torch.empty((1, H, W), dtype=torch.float32).pin_memory()

    def step_device(i):
        a, b = dev_frames[i % nframes]
        feats = net(a, None, b)
        labels, _ = MS.cluster_fields(feats, M, KAPPA, ITERS, [firsts[i]])
        if world > 1:
            dist.all_gather_into_tensor(gathered, labels.view(-1))
        return labels

    def step_e2e(i):
        a, b = pin_frames[i % nframes]
        ad = a.to(dev, non_blocking=True)
        bd = b.to(dev, non_blocking=True)
        feats = net(ad, None, bd)
        labels, _ = MS.cluster_fields(feats, M, KAPPA, ITERS, [firsts[i]])
        if world > 1:
            dist.all_gather_into_tensor(gathered, labels.view(-1))
        out_pin.copy_(labels.view(1, H, W).to(torch.float32), non_blocking=True)   # float32 CPU labels: the reference API
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
    l0 = lib.uoc_launch_count()
    ms_dev = timed(step_device, steps, warmup)
    launches = int(lib.uoc_launch_count() - l0)
    ms_e2e = timed(step_e2e, steps, warmup + steps)

    # ---- pipelined throughput: --depth frames in flight on separate streams --------------------------------
    from unseenobjectclustering_b200.pipeline import FramePipeline
    pipe = FramePipeline(net, H, W, depth=max(1, args.depth), num_seeds=M, kappa=KAPPA, max_iters=ITERS, device=dev,
                         frames_per_slot=B)

    host_ms = [0.0]

    # raw frames (uint8 BGR + uint16 depth as cv2.imread returns them), inputs built on the device (input_prep.py)
    import numpy as np
    rng = np.random.RandomState(7 + rank)
    raw_frames = [(torch.from_numpy(rng.randint(0, 256, (H, W, 3)).astype(np.uint8)).pin_memory(),
                   torch.from_numpy(rng.randint(300, 1500, (H, W)).astype(np.int16)).pin_memory()) for _ in range(nframes)]
    camera = {"fx": 612.937, "fy": 613.173, "x_offset": 322.549, "y_offset": 248.158}   # data/demo/camera_params.json

    def timed_pipe(resident, count, base, raw=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        for sl in pipe.slots:
            sl.stream.wait_event(start)
        ends = []
        host_t0 = time.perf_counter()
        for i in range(count * B):                          # count steps of B frames
            sl = pipe.slots[pipe.next % len(pipe.slots)]
            if sl.fill == 0:
                if sl.busy:
                    pipe.collect_one()
                with torch.cuda.stream(sl.stream):
                    flush.zero_()                           # L2 flush before every step, INSIDE the timed region
            if raw:
                a, b = raw_frames[(base + i) % nframes]
                pipe.submit_raw(a, b, camera, firsts[base + i])
            else:
                a, b = (dev_frames if resident else pin_frames)[(base + i) % nframes]
                pipe.submit(a, b, firsts[base + i], resident=resident)
            if sl.fill != 0:
                continue                                    # the slot's batch is not complete yet
            if world > 1:
                with torch.cuda.stream(sl.stream):
                    dist.all_gather_into_tensor(gathered_b, sl.labels.view(-1))
            e = torch.cuda.Event(enable_timing=True)
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
        timed_pipe(True, 2, i)
        timed_pipe(False, 2, i)
        timed_pipe(False, 2, i, raw=True)
    l0 = lib.uoc_launch_count()
    ms_pipe_dev = timed_pipe(True, steps, warmup)
    host_enqueue_ms = host_ms[0]
    launches_pipe = int(lib.uoc_launch_count() - l0)
    ms_pipe_e2e = timed_pipe(False, steps, warmup + steps)
    ms_pipe_raw = timed_pipe(False, steps, warmup + 2 * steps, raw=True)

    # ---- stage split (same kernels through the stage entry points), rank-local, for the roofline ----
    import ctypes
    ws = MS._workspace(dev, lib.uoc_meanshift_workspace_bytes(1, n, D, M))
    sel = torch.empty((1, M), dtype=torch.int64, device=dev)
    Z = torch.empty((1, M, D), dtype=torch.float32, device=dev)
    sl = torch.empty((1, M), dtype=torch.int32, device=dev)
    nu = torch.empty((1,), dtype=torch.int32, device=dev)
    lab = torch.empty((1, n), dtype=torch.int32, device=dev)
    sp = _lib.stream_ptr(dev)
    stage_ms = {"backbone": 0.0, "fps": 0.0, "loop": 0.0, "label_assign": 0.0}
    reps = max(3, min(steps, 10))
    for rep in range(reps + 1):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        a, b = dev_frames[rep % nframes]
        torch.cuda._sleep(8_000_000)      # ~4 ms of GPU spin: the host enqueues the backbone's ~45 launches meanwhile (else host-bound)
        ev[0].record()
        feats = net(a, None, b)
        xb = MS._lookup_bf16(feats)
        ev[1].record()
        first = (ctypes.c_int64 * 1)(firsts[rep])
        _lib.check(lib.uoc_select_seeds(_lib.ptr(feats), D * n, n, _lib.ptr(xb), 1, n, D, M, ctypes.cast(first, ctypes.c_void_p),
                                        _lib.ptr(sel), _lib.ptr(Z), _lib.ptr(ws), ws.numel(), 0, sp), "select_seeds")
        ev[2].record()
        _lib.check(lib.uoc_hill_climb(_lib.ptr(feats), D * n, n, _lib.ptr(xb), 1, n, D, M, KAPPA, ITERS, _lib.ptr(Z),
                                      _lib.ptr(ws), ws.numel(), 0, sp), "hill_climb")
        ev[3].record()
        _lib.check(lib.uoc_label_seeds(_lib.ptr(Z), 1, M, D, 0.04, _lib.ptr(sl), _lib.ptr(nu), sp), "label_seeds")
        _lib.check(lib.uoc_assign_labels(_lib.ptr(feats), D * n, n, _lib.ptr(xb), 1, n, D, M, _lib.ptr(Z), _lib.ptr(sl), _lib.ptr(nu),
                                         _lib.ptr(lab), _lib.ptr(ws), ws.numel(), sp), "assign_labels")
        ev[4].record()
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for k, (x, y) in zip(("backbone", "fps", "loop", "label_assign"), zip(ev[:-1], ev[1:])):
            stage_ms[k] += x.elapsed_time(y) / reps
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[
            "meanshift_tc_persistent_kernel<64>" if os.environ.get("UOC_LOOP_PERSISTENT", "1") != "0" else "meanshift_tc_kernel<64>"
        ]["dram_bytes_per_launch"]
    except Exception:
        pass
    persistent = os.environ.get("UOC_LOOP_PERSISTENT", "1") != "0"
    if persistent:
        # ONE launch streams the bf16 pixel-major copy ITERS times (one pass per mean-shift update)
        t_launch_s = stage_ms["loop"] * 1e-3
        bytes_actual = ITERS * n * D * 2
        kname = "meanshift_tc_persistent_kernel<64>, one launch = all %d mean-shift updates" % ITERS
    else:
        t_launch_s = stage_ms["loop"] * 1e-3 / ITERS               # one update = tcgen05 kernel + reduce/normalise kernel
        bytes_actual = n * D * 2                                   # bf16 pixel-major copy streamed per update
        kname = "meanshift_tc_kernel<64> (+reduce_normalize_kernel), per mean-shift update"
    achieved = bytes_actual / t_launch_s / 1e9 if t_launch_s > 0 else 0.0
    line = {
        "metric": METRIC, "value": world * B * steps / (ms_pipe_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_pipe_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "steps_in_flight": max(1, args.depth),
                   "l2": "flushed before every step (256 MiB memset on the step's stream, inside the timed region)",
                   "timing": "CUDA events on the launching streams (start -> last frame done), max over ranks",
                   "multi_gpu": "frames sharded, one NCCL all-gather of label maps per step" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": world * B * steps / (ms_pipe_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 2 * 3 * H * W * 4,
                "d2h_bytes_per_step": B * H * W * 4, "ms_per_step": ms_pipe_e2e / steps},
        # same, from RAW host frames (uint8 BGR + uint16 depth): the reference's read_sample arithmetic runs on the device
        "e2e_raw_inputs": {"value": world * B * steps / (ms_pipe_raw * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * H * W * 5,
                           "d2h_bytes_per_step": B * H * W * 4, "ms_per_step": ms_pipe_raw / steps},
        "serial": {"value": world * steps / (ms_dev * 1e-3), "ms_per_step": ms_dev / steps, "e2e_value": world * steps / (ms_e2e * 1e-3),
                   "e2e_ms_per_step": ms_e2e / steps, "note": "one frame at a time, L2 flushed (untimed) between frames"},
        # the pipelined region replays CUDA graphs (not visible to the library's launch counter): the same kernels as the
        # eager serial pass, whose launches were counted
        "gpu_launches": launches, "gpu_launches_eager_in_pipeline": launches_pipe,
        "host_loop_ms_per_step": round(host_enqueue_ms, 3),
        "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "roofline": {"kernel": kname,
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic,
                     "bytes_per_launch": bytes_actual, "fp32_equivalent_GBps": achieved * 2.0,
                     "note": "algorithmic bytes = n*d*2 per mean-shift update (bf16 copy actually streamed); fp32-equivalent (n*d*4) is 2x"},
    }
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        threads = best_cpu_threads(12.0)
        fps, cores, per = cpu_reference_frames(3, 1, threads)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "3 timed full frames (oracle backbone + stage-1 clustering) after 1 warm-up, best torch "
                                          "thread count of a bounded probe, %.1f s in total" % (time.perf_counter() - t0)}
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
    ap.add_argument("--depth", type=int, default=3, help="steps in flight per GPU (1 = strictly serial)")
    ap.add_argument("--batch", type=int, default=4, help="frames per GPU per step: they go through every kernel together")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
