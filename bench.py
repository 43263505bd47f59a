#!/usr/bin/env python
"""bench.py -- A-scan throughput of the raw -> B-scan hot path on the reference's headline workload.

  python bench.py --gpus N --steps K --warmup W            # our arm (liboctb200.so on B200)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's own CPU path on the host cores

Workload (BASELINE.json configs[1], the config the metric is quoted on): the 1024 x 512 x 256 12-bit test
volume of the reference's published benchmark (performance/v180/.../20250504_octproz_settings.ini): cubic
k-linearisation + dispersion + Hann window + FPN (1 B-scan, determined once) + log scaling.  The reference's
dataset is an external download, so the raw buffer is synthetic with that geometry (octproz_b200/synth.py).
A "step" = one pass of the hot path over one raw buffer (one volume = 131072 A-scans) per GPU.

One JSON line on stdout (rank 0):
  value      device-resident rate (raw already in HBM), CUDA events on the library's compute stream, max over ranks.
             N > 1: WEAK scaling -- every rank processes its own 256-B-scan buffer (a slab of an N-times larger volume); the
             en-face slice of the whole volume is gathered AND consumed (wait + copy-out + acknowledge) on every rank, every step.
  strong     N > 1: the ONE 256-B-scan volume split 256/N B-scans per rank (BASELINE.json configs[4]), same gather, every step.
  e2e        through the reference-facing call octCudaPipeline(host buffer): pinned H2D + D2H of the converted output inside
             the timed region (the reference's stream-to-host path).
  roofline   algorithmic bytes / event-timed launches of the dominant kernel / MEASURED_PEAKS.json HBM.
  cpu_baseline, ref_cuda   the reference's CPU path on the host cores and the reference's unmodified CUDA build on the same
             GPU (oracle/_ref, separate process), timed beside ours on the same buffers (N = 1, rank 0).
  secondary  BASELINE configs 2 / 3 / 4 (cuFFT chain, split chain, 2048 x 1024 x 512 16-bit with FPN + flip + sinusoidal), device-resident.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (samplesPerLine, ascansPerBscan, bscansPerBuffer, bitDepth)
    "1024x512x256-12bit": (1024, 512, 256, 12),
    "2048x1024x128-16bit": (2048, 1024, 128, 16),
    # the reference's DEFAULT line length (octproz/default/settings.ini:62) at the BASELINE volume shape: FUSED = the shared-memory kernel
    "1664x512x256-12bit": (1664, 512, 256, 12),
    # BASELINE.json configs[3]: "2048-sample A-scan x 1024 x 512 16-bit, full pipeline incl. FPN + sinusoidal correction" (2 GiB raw per buffer)
    "2048x1024x512-16bit-config4": (2048, 1024, 512, 16),
}
# parameter overrides on top of the benchmark INI settings, per workload
WORKLOAD_PARAMS = {"2048x1024x512-16bit-config4": dict(bscanFlip=True, sinusoidalScanCorrection=True)}
DEFAULT_WORKLOAD = "1024x512x256-12bit"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
UNIT = "MHz (1e6 A-scans/s)"
METRIC = "A-scan rate (raw -> B-scan hot path)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """the cores this process may run on -- NOT omp_get_max_threads(): torch.distributed.run exports OMP_NUM_THREADS=1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def workload_params(name):
    from octproz_b200 import benchmark_params
    n, a, b, bits = WORKLOADS[name]
    q = benchmark_params(n, a, b, bits)
    for k, v in WORKLOAD_PARAMS.get(name, {}).items():
        setattr(q, k, v)
    q.update_all_curves()
    return q


def make_config(args, world):
    """the `config` object -- built by ONE function so that both arms print the identical dict"""
    n, a, b, bits = WORKLOADS[args.workload]
    extra = " + B-scan flip + sinusoidal scan correction" if WORKLOAD_PARAMS.get(args.workload, {}).get("sinusoidalScanCorrection") else ""
    return {"workload": f"{args.workload} volume, benchmark INI settings (cubic k-lin + dispersion + Hann + FPN once + log){extra}, u16 container, synthetic",
            "samples_per_ascan": n, "ascans_per_bscan": a, "bscans_per_buffer": b, "bit_depth": bits,
            "step": f"one {b}-B-scan raw buffer ({a * b} A-scans) per GPU",
            "l2": f"two alternating {n * a * b * 2 >> 20} MiB inputs and {(n // 2) * a * b * 4 >> 20} MiB outputs per step: larger than the 126 MB L2 "
                  "(GPU arm); the CPU arm processes the same full buffer per step",
            "parallelism": f"b-scan sharding x{world}"}


def make_raw(q, bscans_unique=8, seed_offset=0):
    from octproz_b200 import synth
    small = synth.make_volume(q.samplesPerLine, q.ascansPerBscan, bscans_unique, q.bitDepth, resample=q.resampleCurve,
                              dispersion=q.dispersionCurve, b_offset=seed_offset)
    reps = (q.bscansPerBuffer + bscans_unique - 1) // bscans_unique
    return np.ascontiguousarray(np.tile(small, (reps, 1, 1))[: q.bscansPerBuffer])


# ------------------------------------------------------------------------------------------------ CPU reference
_BEST_THREADS = {}


def best_cpu_threads(q, ncores):
    """the reference's CPU path allocates per A-scan (processor.tpp:256-317); with very many threads the allocator and
    page-fault traffic can make it slower, so pick the fastest of a few thread counts on a small sample (reported in `cores`)."""
    key = (q.samplesPerLine, q.ascansPerBscan, ncores)
    if key in _BEST_THREADS:
        return _BEST_THREADS[key]
    cands = sorted({t for t in (ncores, ncores // 2, ncores // 4, 32, 16, 8) if 1 <= t <= ncores}, reverse=True)
    best, best_rate = 1, 0.0
    for t in cands:
        r = cpu_reference_run(q, t, max(8, min(2 * t, 64)), calibrate=False)
        if r["mhz"] > best_rate:
            best, best_rate = r["threads"], r["mhz"]
    _BEST_THREADS[key] = best
    return best


def cpu_reference_run(q, threads, bscans, repeats=1, calibrate=True, raw=None):
    """time the reference's own CPU path (oracle/_ref/libref_cpu.so, FFTW-API substitute) on `bscans` B-scans.  The thread count is
    passed explicitly (OpenMP num_threads clause in oracle/ref_drivers/ref_cpu.cpp), so OMP_NUM_THREADS does not limit it."""
    from oracle import oracle as orc
    kind = "reference"
    if orc.have_ref("libref_cpu.so"):
        rc = orc.RefCpu()
        threads = max(1, threads)
        if calibrate:
            threads = best_cpu_threads(q, threads)
        run = lambda r: rc.process(q, r, threads=threads)
    else:
        kind, threads = "port", 1
        run = lambda r: orc.process(q, r, precision=32)
    if raw is None:
        q2 = copy.copy(q); q2.bscansPerBuffer = bscans
        raw = make_raw(q2, bscans_unique=min(8, bscans))
    else:
        raw = raw[:bscans]
    run(raw[: max(1, min(bscans, threads))])   # warm caches / thread pool / per-thread arenas
    t0 = time.perf_counter()
    for _ in range(repeats):
        run(raw)
    dt = (time.perf_counter() - t0) / repeats
    ascans = bscans * q.ascansPerBscan
    return {"seconds": dt, "ascans": ascans, "mhz": ascans / dt / 1e6, "threads": threads, "kind": kind,
            "sample": f"{bscans} B-scans ({ascans} A-scans) of the workload per step, reference CPU path "
                      f"(processor.tpp, FFTW-API substitute), {threads} thread(s)"}


def cpu_single_thread(q, bscans=4):
    """the reference's CPU path "as shipped" runs on one thread; a small sample is enough for a rate"""
    try:
        r = cpu_reference_run(q, 1, bscans, repeats=1, calibrate=False)
        return {"value": r["mhz"], "unit": UNIT, "sample": r["sample"]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def reference_arm(args, q, config):
    """--impl reference: the reference's CPU implementation of the path on every host core, the FULL buffer per step"""
    n, a, b, bits = WORKLOADS[args.workload]
    ncores = host_threads()
    bscans = args.cpu_bscans or b
    raw = make_raw(q)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_run(q, ncores, bscans, raw=raw)
    t = [cpu_reference_run(q, ncores, bscans, raw=raw) for _ in range(args.steps)]
    sec = float(np.mean([x["seconds"] for x in t])); mhz = t[0]["ascans"] / sec / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": mhz, "unit": UNIT,
            "volumes_per_s": mhz * 1e6 / (a * b), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": mhz, "unit": UNIT, "cores": t[0]["threads"], "kind": t[0]["kind"], "sample": t[0]["sample"],
                             "single_thread": cpu_single_thread(q)},
            "e2e": {"value": mhz, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": ncores}
    print(json.dumps(line), flush=True)
    return 0


def ref_cuda_leg(workload, steps):
    """the reference's UNMODIFIED cuda_code.cu (oracle/_ref/libref_cuda.so) on this GPU, in a process of its own (file-scope globals,
    exit() on CUDA errors): device-resident and end-to-end on the same synthetic buffers -- BASELINE.md's "number to beat"."""
    exe = os.path.join(ROOT, "tools", "ref_cuda_bench.py")
    if not (os.path.exists(exe) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so"))):
        return {"unavailable": "oracle/_ref/libref_cuda.so is not built (it needs /root/reference at build time)"}
    try:
        env = dict(os.environ); env.pop("OMP_NUM_THREADS", None)
        r = subprocess.run([sys.executable, exe, workload, str(max(3, min(steps, 20)))], capture_output=True, text=True, timeout=600, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"no result (rc {r.returncode}): {(r.stderr or r.stdout)[-300:]}"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


# ------------------------------------------------------------------------------------------------ our arm
class Rig:
    """one pipeline + its device inputs for a geometry; the pieces every measurement below shares"""

    def __init__(self, q, mode, local, rank, world, raw_np, bscan_base=0, flags=0, packing=None, total_bscans=0):
        import torch
        from octproz_b200 import OctPipeline, _lib
        self.torch, self.q, self.world, self.rank = torch, copy.deepcopy(q), world, rank
        kw = {} if packing is None else {"input_packing": packing}
        if total_bscans:
            kw["bscans_in_unsharded_buffer"] = total_bscans
        self.p = OctPipeline(fft_mode=mode, device=local, bscan_index_base=bscan_base, flags=flags, **kw)
        if not self.p.initializeCuda(None, None, self.q):
            raise SystemExit("initializeCuda failed: " + getattr(self.p, "_create_error", ""))
        self.h_raw = [torch.from_numpy(x).pin_memory() for x in raw_np]
        self.d_raw = [x.cuda(non_blocking=False) for x in self.h_raw]
        self.gather = None

    def connect_gather(self, dist, global_lines, line_offset, frame=100):
        """peer-memory en-face gather (octb200_enface_gather_*): exchange the IPC handles once, then every process call gathers"""
        torch, p = self.torch, self.p
        try:
            mine = p.enface_gather_init(self.rank, self.world, global_lines, line_offset)
            t = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t)
            p.enface_gather_connect(b"".join(x.cpu().numpy().tobytes() for x in parts))
            ok = torch.ones(1, device="cuda")
        except Exception as e:  # noqa: BLE001
            print(f"rank {self.rank}: peer-memory en-face gather unavailable ({e}); using NCCL", file=sys.stderr, flush=True)
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.gather = "p2p" if float(ok.item()) > 0 else "nccl"
        dist.barrier()
        if self.gather == "p2p":
            p.enface_gather_auto(True, frame, 1, 0)
        return self.gather

    def gather_health(self, dist):
        """after a few warm-up steps: the device-side time-out counters of the peer gather, worst rank (0 / 0 in a healthy run).  If any
        rank timed out, every rank drops to the NCCL gather together -- a stalled protocol costs 10 s per wait and would turn the timed
        region into a measurement of the time-out -- and the JSON line says so."""
        if self.gather != "p2p":
            return None
        torch = self.torch
        try:
            st = self.p.enface_gather_status()
            worst = torch.tensor([float(st["ack_timeouts"]), float(st["arrival_timeouts"])], device="cuda")
        except Exception as e:  # noqa: BLE001
            print(f"rank {self.rank}: enface_gather_status failed ({e})", file=sys.stderr, flush=True)
            worst = torch.zeros(2, device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        health = {"ack_timeouts_worst_rank": int(worst[0].item()), "arrival_timeouts_worst_rank": int(worst[1].item())}
        if health["ack_timeouts_worst_rank"] or health["arrival_timeouts_worst_rank"]:
            self.p.enface_gather_auto(False)
            self.gather = "nccl"
            health["fallback"] = "peer gather timed out during warm-up: NCCL all-gather used for the timed region"
        return health

    def close(self, dist=None):
        if dist is not None:
            self.p.sync(); self.torch.cuda.synchronize(); dist.barrier()       # peers have stopped writing into this rank's window
            if self.gather == "p2p":
                self.p.enface_gather_close()
        self.p.cleanupCuda()


def timed_resident(rig, steps, warmup, dist, step_fn):
    """W warm-up steps, then K steps between two events on the library's compute stream; max over ranks"""
    torch, p = rig.torch, rig.p

    def sync_all():
        p.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()
    for i in range(max(3, warmup)):
        step_fn(i)
    sync_all()
    l0 = p.launch_count()
    # a ~1 ms spin kernel in front of the first event: the host enqueues the timed steps while it runs, so the timed region starts
    # with work queued and measures the device, not the launch latency / scheduling hiccups of a (shared) host
    try:
        with torch.cuda.stream(torch.cuda.ExternalStream(int(p._lib.octb200_compute_stream(p.handle)))):
            torch.cuda._sleep(2_000_000)
    except Exception:  # noqa: BLE001
        pass
    p.event_record(0)
    for i in range(steps):
        step_fn(i)
    if rig.gather == "p2p":
        p.enface_gather_wait()             # (every frame of the region was consumed inside it: the consumer kernels run in stream order)
    p.event_record(1)
    ms_total = p.event_elapsed_ms(0, 1)
    sync_all()
    launches = p.launch_count() - l0
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
    return ms_total / steps, launches


def gather_check(rig, dist, a_lines_local, frame=100):
    """once, outside every timed region: the frame assembled by the library's peer-memory gather on THIS rank against the en-face
    slices extracted per rank and all-gathered by NCCL"""
    import ctypes as C
    torch, p = rig.torch, rig.p
    world = rig.world
    p.process_device(rig.d_raw[0])
    ptr = p.enface_gather_wait()
    mine = torch.empty(a_lines_local, dtype=torch.float32, device="cuda")
    p.changeDisplayedEnFaceFrame(frame, 1, 0, mine)
    p.sync(); torch.cuda.synchronize()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    # updateDisplayedEnFaceViewFrame writes disp[(E-1) - i] (cuda_code.cu:909): the frame of the whole volume is the per-rank frames in reverse rank order
    want = torch.cat(list(reversed(parts))).cpu().numpy()
    got = np.empty(world * a_lines_local, np.float32)
    cudart = C.CDLL("libcudart.so.12") if False else None  # noqa: F841  (torch moves the bytes below; no direct cudart use)
    t = torch.empty(world * a_lines_local, dtype=torch.float32, device="cuda")
    from octproz_b200.pipeline import device_copy
    device_copy(t, ptr, t.numel() * 4)
    got = t.cpu().numpy()
    bad = int(np.sum(got != want))
    ok = torch.tensor([1.0 if bad == 0 else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"ok": bool(ok.item() > 0), "mismatches_on_rank0": bad, "elements": int(got.size),
            "checked_against": "per-rank en-face extraction + ncclAllGather, bit-exact, every rank"}


def link_probe(rig, conv_bytes, dist):
    """what the host link of this box can do (context for e2e): pinned H2D of one raw buffer alone, with a D2H of the converted-output
    size running the other way, and -- N > 1 -- with every rank copying at the same time (aggregate)"""
    torch = rig.torch
    try:
        bytes_in = rig.h_raw[0].numel() * rig.h_raw[0].element_size()
        d_probe = torch.empty_like(rig.d_raw[0]); d_conv = torch.empty(conv_bytes, dtype=torch.uint8, device="cuda")
        h_conv = torch.empty(conv_bytes, dtype=torch.uint8).pin_memory()
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def h2d_rate(with_d2h, reps=8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s_up):
                e0.record()
                for i in range(reps):
                    d_probe.copy_(rig.h_raw[i & 1], non_blocking=True)
                e1.record()
            if with_d2h:
                with torch.cuda.stream(s_dn):
                    for i in range(2 * reps):
                        h_conv.copy_(d_conv, non_blocking=True)
            torch.cuda.synchronize()
            return bytes_in * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_rate(False, 2)
        alone, duplex = h2d_rate(False), h2d_rate(True)
        link = {"h2d_gbs_alone": alone, "h2d_gbs_with_concurrent_d2h": duplex}
        if dist is not None:
            dist.barrier()
            mine = h2d_rate(True)
            t = torch.tensor([mine], device="cuda")
            s = t.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
            m = t.clone(); dist.all_reduce(m, op=dist.ReduceOp.MIN)
            link.update(all_ranks_concurrent_h2d_gbs_sum=float(s.item()), all_ranks_concurrent_h2d_gbs_min=float(m.item()),
                        note="every rank copies its own pinned buffer to its own GPU at the same time, D2H running the other way")
        del d_probe, d_conv, h_conv
        return link
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def e2e_leg(rig, steps, warmup, dist, conv_bytes, hook=None):
    """octCudaPipeline(host buffer) with streaming to the host on: H2D of the raw buffer, the chain, floatToOutput, D2H of the converted
    buffer -- all inside the timed region (host wall clock between device synchronisations, max over ranks)"""
    torch, p = rig.torch, rig.p
    h_stream = [np.zeros(conv_bytes, np.uint8) for _ in range(2)]   # plain host memory; the library pins it like the reference (cuda_code.cu:661)
    p.sync()
    rig.q.streamToHost = True
    p.cuda_registerStreamingBuffers(h_stream[0], h_stream[1], conv_bytes)
    for i in range(max(3, warmup)):
        p.octCudaPipeline(rig.h_raw[i & 1].numpy())
    p.sync(); torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    n_steps = max(1, min(steps, 200))
    l0 = p.launch_count()
    t0 = time.perf_counter()
    for i in range(n_steps):
        p.octCudaPipeline(rig.h_raw[i & 1].numpy())
        if hook:
            hook(i)
    p.sync(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    launches = (p.launch_count() - l0) / n_steps
    checksum = int(h_stream[0][:4096].view(np.uint16).sum()) + int(h_stream[1][:4096].view(np.uint16).sum())
    p.cuda_unregisterStreamingBuffers()
    rig.q.streamToHost = False
    return dt, n_steps, launches, checksum


def secondary_resident(name, mode_name, local, steps):
    """device-resident rate of another BASELINE config on this GPU (N = 1): its own pipeline, two alternating inputs"""
    import torch
    from octproz_b200 import _lib
    mode = {"fused": _lib.FFT_FUSED, "split": _lib.FFT_SPLIT, "cufft": _lib.FFT_CUFFT}[mode_name]
    n, a, b, bits = WORKLOADS[name]
    try:
        q = workload_params(name)
        raw = make_raw(q)
        rig = Rig(q, mode, local, 0, 1, [raw, raw[::-1].copy()])
        rig.p.process_device(rig.d_raw[0]); rig.p.sync()
        k = max(3, min(steps, 50))
        ms, launches = timed_resident(rig, k, 3, None, lambda i: rig.p.process_device(rig.d_raw[i & 1]))
        kern_ms = rig.p.time_kernel(rig.d_raw[1], 10)
        per_sample = {"fused": 4, "split": 10, "cufft": 10}[mode_name]
        peak, _ = hbm_peak()
        out = {"workload": name, "mode": mode_name, "value": a * b / (ms * 1e3), "unit": UNIT, "ms_per_step": ms, "steps": k,
               "gpu_launches_per_step": launches / k,
               "dominant_kernel": {"name": {"fused": "oct_fused_kernel" if n in (1024, 2048) else "oct_generic_kernel", "split": "oct_fused_kernel<SRC_CPLX>", "cufft": "oct_pre_kernel"}[mode_name],
                                   "kernel_ms": kern_ms, "algorithmic_bytes_per_sample": per_sample,
                                   "frac_of_hbm_peak": a * b * n * per_sample / (kern_ms * 1e-3) / 1e9 / peak}}
        rig.close()
        del rig
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # noqa: BLE001
        return {"workload": name, "mode": mode_name, "error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--mode", default="fused", choices=["fused", "split", "cufft"])
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"],
                    help="N > 1: `value` is always the weak-scaling rate; `strong` adds the one-volume-split-N-ways curve")
    ap.add_argument("--enface", default="p2p", choices=["p2p", "nccl"], help="multi-GPU en-face gather: own peer-memory kernel or NCCL")
    ap.add_argument("--no-packed", action="store_true", help="skip the 12-bit packed-input extension measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configs (cuFFT / split chains, config 4) and the reference CUDA leg")
    ap.add_argument("--cpu-bscans", type=int, default=0, help="B-scans per step of the CPU reference (0 = the full buffer)")
    ap.add_argument("--no-numa", action="store_true", help="do not move the host thread to the GPU-local CPUs before pinning the host buffers")
    ap.add_argument("--separate-conversion", action="store_true",
                    help="end-to-end leg: floatToOutput as its own pass (the reference's order) instead of folded into the fused kernel's epilogue")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n, a, b, bits = WORKLOADS[args.workload]
    q = workload_params(args.workload)
    config = make_config(args, world)
    ascans_per_step = a * b

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, q, config)

    # ------------------------------------------------------------------ our arm (B200)
    # stdout carries exactly ONE line, the JSON: libraries that print there (NCCL's version banner at the first communicator) are
    # sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from octproz_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = {"fused": _lib.FFT_FUSED, "split": _lib.FFT_SPLIT, "cufft": _lib.FFT_CUFFT}[args.mode]

    # host buffers are pinned from a thread on the GPU-local CPUs: first touch puts them on the GPU's NUMA node (octproz_b200/hostmem.py)
    from octproz_b200.hostmem import local_affinity
    aff = local_affinity(local, cpus=set() if args.no_numa else None)
    host_numa = aff.__enter__()

    raw_np = [make_raw(q, seed_offset=16 * rank), make_raw(q, seed_offset=16 * rank + 8)]   # two distinct inputs per rank, each larger than L2
    bytes_in = raw_np[0].nbytes
    conv_bytes = (n // 2) * a * b * 2
    rig = Rig(q, mode, local, rank, world, raw_np, bscan_base=(rank * b) % 2,
              flags=_lib.FLAG_SEPARATE_CONVERSION if args.separate_conversion else 0)
    p = rig.p
    enface = torch.empty(a * b, dtype=torch.float32, device="cuda")
    gathered = torch.empty(world * a * b, dtype=torch.float32, device="cuda") if world > 1 else None
    stream = torch.cuda.ExternalStream(int(p._lib.octb200_compute_stream(p.handle)), device=torch.device("cuda", local))

    def fpn_share(r):
        # FPN line of the first buffer: rank 0 determines, everyone uses it (8 KB broadcast)
        if dist is None:
            return
        ml = torch.from_numpy(r.p.fpn_mean_line()).cuda()
        dist.broadcast(ml, 0)
        r.p.set_fpn_mean_line(ml.cpu().numpy())

    # en-face slice of the N-times larger volume, every step: the fused kernel's epilogue stores each rank's values straight into every
    # rank's frame window over NVLink peer memory (octb200_enface_gather_*); the consuming side waits for all slabs, copies the frame out
    # and acknowledges (flow control) -- also every step.  `--enface nccl`: extraction kernel + ncclAllGather instead (baseline).
    gather_impl = None
    if dist is not None:
        gather_impl = "nccl" if args.enface == "nccl" else rig.connect_gather(dist, world * a * b, rank * a * b)

    def step_resident(i, r=rig, frame=enface, out=gathered):
        # p2p: the process call itself gathers (peer stores from the kernel's epilogue) and enqueues the consumer kernel of this frame
        # (wait for every rank's slab, copy the frame out, acknowledge) behind it
        r.p.process_device(r.d_raw[i & 1])
        if r.gather == "nccl":
            r.p.changeDisplayedEnFaceFrame(100, 1, 0, frame)
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(out, frame)

    # warm-up (includes LUT build, FPN determination, cuFFT plan if any)
    p.process_device(rig.d_raw[0]); p.sync(); fpn_share(rig)

    ghealth = None
    if dist is not None:
        if args.enface == "nccl":
            rig.gather = "nccl"
        for i in range(4):
            step_resident(i)
        rig.p.sync(); torch.cuda.synchronize()
        ghealth = rig.gather_health(dist)
        gather_impl = rig.gather
    sampler = ClockSampler(local); sampler.start()     # samples run until the end of the end-to-end region
    ms_step, launches = timed_resident(rig, args.steps, args.warmup, dist, step_resident)
    value = world * ascans_per_step / (ms_step * 1e3)   # MHz

    gcheck = None
    if gather_impl == "p2p":
        gcheck = gather_check(rig, dist, a * b)

    # ---- dominant kernel alone (roofline): events around back-to-back launches of the fused kernel ----
    kern_ms = p.time_kernel(rig.d_raw[1], 20)
    # algorithmic bytes of the timed kernel per raw sample: fused = 2 B in + 2 B out; the split path's FFT kernel reads
    # the float2 FFT input (8 B) and writes 2 B; the cuFFT path's pre kernel reads 2 B and writes 8 B
    alg_bytes = ascans_per_step * n * {"fused": 4, "split": 10, "cufft": 10}[args.mode]
    peak, peak_src = hbm_peak()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.mode}:{args.workload}")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": {"fused": "oct_fused_kernel", "split": "oct_fused_kernel<SRC_CPLX>", "cufft": "oct_pre_kernel"}[args.mode],
                "achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / peak,
                "traffic": traffic, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}

    # ---- end to end through octCudaPipeline(host buffer) ----
    e2e_hook = None
    e2e_s, e2e_steps, e2e_launches, checksum = e2e_leg(rig, args.steps, args.warmup, dist, conv_bytes, e2e_hook)
    clocks = sampler.stop()
    e2e_mhz = world * ascans_per_step * e2e_steps / e2e_s / 1e6
    link = link_probe(rig, conv_bytes, dist)
    if "error" not in link:
        rate = bytes_in * e2e_steps / e2e_s / 1e9
        link.update(e2e_h2d_gbs_per_gpu=rate, e2e_frac_of_duplex_h2d=rate / link["h2d_gbs_with_concurrent_d2h"])
        if "all_ranks_concurrent_h2d_gbs_min" in link:
            link["e2e_frac_of_all_ranks_concurrent_link"] = rate * world / link["all_ranks_concurrent_h2d_gbs_sum"]

    # ---- strong scaling (N > 1): the ONE volume split B/N B-scans per rank; FPN line from rank 0's shard; gather + consume every step ----
    strong = None
    if dist is not None and args.scaling in ("strong", "both") and args.mode == "fused" and b % world == 0:
        bs = b // world
        qs = copy.deepcopy(q); qs.bscansPerBuffer = bs
        # every rank generates the SAME two volumes and keeps its own B-scans
        vols = [make_raw(q, seed_offset=0), make_raw(q, seed_offset=8)]
        shard = [np.ascontiguousarray(v[rank * bs:(rank + 1) * bs]) for v in vols]
        rs = Rig(qs, mode, local, rank, world, shard, bscan_base=rank * bs, total_bscans=b)
        g2 = "nccl" if args.enface == "nccl" else rs.connect_gather(dist, a * b, rank * a * bs)
        en_s = torch.empty(a * bs, dtype=torch.float32, device="cuda"); ga_s = torch.empty(a * b, dtype=torch.float32, device="cuda")

        strong_call, strong_h, strong_ptr = rs.p._lib.octb200_process_device, rs.p.handle, [int(t.data_ptr()) for t in rs.d_raw]

        def step_strong(i):
            # the bare C-ABI call (parameters do not change between buffers): at N = 8 a buffer is ~30 us of GPU time, the Python
            # parameter marshalling of OctPipeline.process_device alone would be as long
            if strong_call(strong_h, strong_ptr[i & 1]) != 0:
                raise RuntimeError("octb200_process_device failed")
            if rs.gather != "p2p":
                rs.p.changeDisplayedEnFaceFrame(100, 1, 0, en_s)
                with torch.cuda.stream(torch.cuda.ExternalStream(int(rs.p._lib.octb200_compute_stream(rs.p.handle)), device=torch.device("cuda", local))):
                    dist.all_gather_into_tensor(ga_s, en_s)
        rs.p.process_device(rs.d_raw[0]); rs.p.sync(); fpn_share(rs)
        if g2 != "p2p":
            rs.gather = "nccl"
        for i in range(4):
            step_strong(i)
        rs.p.sync(); torch.cuda.synchronize()
        health_s = rs.gather_health(dist)
        g2 = rs.gather
        k_strong = max(args.steps, 50)
        ms_s, l_s = timed_resident(rs, k_strong, args.warmup, dist, step_strong)
        sc = gather_check(rs, dist, a * bs) if g2 == "p2p" else None
        kern_s = rs.p.time_kernel(rs.d_raw[1], 50)
        strong = {"scaling": "strong", "value": ascans_per_step / (ms_s * 1e3), "unit": UNIT, "volumes_per_s": 1e3 / ms_s, "ms_per_volume": ms_s,
                  "steps": k_strong, "bscans_per_rank": bs, "gpu_launches_per_step": l_s / k_strong, "fused_kernel_alone_ms": kern_s,
                  "tail_us_per_step": (ms_s - kern_s) * 1e3, "gather": g2, "gather_check": sc, "gather_health": health_s,
                  "note": "each step = the whole 1024x512x256 volume; per-rank inputs (32 MiB at N = 8) alternate between two buffers and stay L2-resident "
                          "at N >= 4 -- the kernel is not HBM-bound, so this does not flatter it"}
        rs.close(dist)
        del rs

    # ---- extension beside the headline (12-bit workload): the same volume delivered 12-bit PACKED (3 bytes per 2 samples,
    #      include/octb200.h OCTB200_PACK_12P; the reference only takes containers).  Reported separately, never as `value` / `e2e`. ----
    packed = None
    if bits == 12 and args.mode == "fused" and not args.no_packed:
        from octproz_b200.packing import pack12
        rp = Rig(q, mode, local, rank, world, [pack12(x) for x in raw_np], bscan_base=(rank * b) % 2, packing=_lib.PACK_12P)
        rp.p.process_device(rp.d_raw[0]); rp.p.sync(); fpn_share(rp)
        n_res = max(3, min(args.steps, 200))
        ms_p, _ = timed_resident(rp, n_res, 5, dist, lambda i: rp.p.process_device(rp.d_raw[i & 1]))
        dt, n_e2e, _, cs = e2e_leg(rp, min(args.steps, 100), 3, dist, conv_bytes)
        packed = {"input": "12-bit packed (Mono12p), extension", "value": world * ascans_per_step / (ms_p * 1e3), "ms_per_step": ms_p,
                  "e2e": {"value": world * ascans_per_step * n_e2e / dt / 1e6, "h2d_bytes_per_step": int(rp.h_raw[0].numel()), "d2h_bytes_per_step": conv_bytes,
                          "ms_per_step": dt * 1e3 / n_e2e, "steps": n_e2e},
                  "unit": UNIT, "checksum": cs}
        rp.close()
        del rp

    aff.__exit__(None, None, None)          # the CPU baseline gets every core again
    rig.close(dist)
    del rig
    torch.cuda.empty_cache()

    # ---- baselines and secondary configs beside it (rank 0, N = 1 only) ----
    cpu = ref_cuda = secondary = None
    if rank == 0 and world == 1:
        ncores = host_threads()
        bscans = args.cpu_bscans or b            # one full volume per repeat
        c = cpu_reference_run(q, ncores, bscans, repeats=4, raw=raw_np[0])
        cpu = {"value": c["mhz"], "unit": UNIT, "cores": c["threads"], "kind": c["kind"], "sample": c["sample"],
               "single_thread": cpu_single_thread(q)}
        if not args.no_secondary:
            secondary = []
            if args.workload == DEFAULT_WORKLOAD and args.mode == "fused":
                secondary.append(secondary_resident(DEFAULT_WORKLOAD, "cufft", local, args.steps))     # BASELINE configs[1]
                secondary.append(secondary_resident(DEFAULT_WORKLOAD, "split", local, args.steps))     # BASELINE configs[2]
                secondary.append(secondary_resident("2048x1024x512-16bit-config4", "fused", local, min(args.steps, 20)))   # BASELINE configs[3]
                secondary.append(secondary_resident("1664x512x256-12bit", "fused", local, args.steps))      # the reference's default line length:
                secondary.append(secondary_resident("1664x512x256-12bit", "cufft", local, args.steps))      # own shared-memory FFT vs the cuFFT chain
            ref_cuda = ref_cuda_leg(args.workload, args.steps)

    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": METRIC, "value": value, "unit": UNIT,
            "volumes_per_s": value * 1e6 / ascans_per_step, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "ours", "config": config, "mode": args.mode,
            "gather": None if world == 1 else {"impl": gather_impl, "every_step": "peer-memory stores from the fused kernel's epilogue + consumer kernel (wait for all slabs, copy out, acknowledge) in stream order"
                                               if gather_impl == "p2p" else "extraction kernel + ncclAllGather", "check": gcheck, "health": ghealth},
            "e2e": {"value": e2e_mhz, "unit": UNIT, "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": conv_bytes,
                    "ms_per_step": e2e_s * 1e3 / e2e_steps, "steps": e2e_steps, "timer": "host wall clock between device synchronisations, max over ranks",
                    "checksum": checksum, "gpu_launches_per_step": e2e_launches,
                    "conversion": "floatToOutput as a separate pass" if args.separate_conversion or args.mode != "fused"
                    else "floatToOutput folded into the fused kernel's epilogue (u16 line written beside the float line)",
                    "host_numa": host_numa, "link": link},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "host_cores": host_threads()}
    if strong is not None:
        line["strong"] = strong
    if packed is not None:
        line["packed12"] = packed
    if ref_cuda is not None:
        line["ref_cuda"] = ref_cuda
    if secondary is not None:
        line["secondary"] = secondary
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
