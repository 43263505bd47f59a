#!/usr/bin/env python
"""bench.py -- A-scan throughput of the raw -> B-scan hot path on the reference's headline workload.

  python bench.py --gpus N --steps K --warmup W            # our arm (liboctb200.so on B200)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's own CPU path on the host cores

Workload (BASELINE.json configs[1], the config the metric is quoted on): the 1024 x 512 x 256 12-bit test
volume of the reference's published benchmark (performance/v180/.../20250504_octproz_settings.ini): cubic
k-linearisation + dispersion + Hann window + FPN (1 B-scan, determined once) + log scaling.  The reference's
dataset is an external download, so the raw buffer is synthetic with that geometry (octproz_b200/synth.py).
A "step" = one pass of the hot path over one raw buffer (one volume = 131072 A-scans) per GPU.

One JSON line on stdout (rank 0).  `value` = device-resident rate (raw already in HBM), `e2e` = through the
reference-facing call octCudaPipeline(host buffer) with the pinned H2D copy and the D2H of the converted
output (the reference's stream-to-host path) inside the timed region.  Multi-GPU: weak scaling, every rank
processes its own buffer (slab of a G-times larger volume) and the en-face slice is gathered inside the timed
region, every step, by the library's own kernel over NVLink peer memory (octb200_enface_gather; `--enface nccl` = the
extraction kernel + ncclAllGather baseline).
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
    # BASELINE.json configs[3]: "2048-sample A-scan x 1024 x 512 16-bit, full pipeline incl. FPN + sinusoidal correction" (2 GiB raw per buffer)
    "2048x1024x512-16bit-config4": (2048, 1024, 512, 16),
}
# parameter overrides on top of the benchmark INI settings, per workload
WORKLOAD_PARAMS = {"2048x1024x512-16bit-config4": dict(bscanFlip=True, sinusoidalScanCorrection=True)}
DEFAULT_WORKLOAD = "1024x512x256-12bit"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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


def make_raw(q, bscans_unique=8, seed_offset=0):
    from octproz_b200 import synth
    small = synth.make_volume(q.samplesPerLine, q.ascansPerBscan, bscans_unique, q.bitDepth, resample=q.resampleCurve,
                              dispersion=q.dispersionCurve, b_offset=seed_offset)
    reps = (q.bscansPerBuffer + bscans_unique - 1) // bscans_unique
    return np.ascontiguousarray(np.tile(small, (reps, 1, 1))[: q.bscansPerBuffer])


_BEST_THREADS = {}


def best_cpu_threads(q, ncores):
    """the reference's CPU path allocates per A-scan (processor.tpp:256-317); with very many threads the allocator and
    page-fault traffic can make it slower, so pick the fastest of a few thread counts on a small sample (reported in `cores`)."""
    key = (q.samplesPerLine, q.ascansPerBscan)
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


def cpu_reference_run(q, threads, bscans, repeats=1, calibrate=True):
    """time the reference's own CPU path (oracle/_ref/libref_cpu.so, FFTW-API substitute) on `bscans` B-scans"""
    from oracle import oracle as orc
    kind = "reference"
    if orc.have_ref("libref_cpu.so"):
        rc = orc.RefCpu()
        threads = min(threads, rc.max_threads) if threads > 0 else rc.max_threads
        if calibrate:
            threads = best_cpu_threads(q, threads)
        run = lambda raw: rc.process(q, raw, threads=threads)
    else:
        kind, threads = "port", 1
        run = lambda raw: orc.process(q, raw, precision=32)
    q2 = copy.copy(q); q2.bscansPerBuffer = bscans
    raw = make_raw(q2, bscans_unique=min(8, bscans))
    run(raw[: max(1, min(bscans, threads))])   # warm caches / thread pool / per-thread arenas
    t0 = time.perf_counter()
    for _ in range(repeats):
        run(raw)
    dt = (time.perf_counter() - t0) / repeats
    ascans = bscans * q.ascansPerBscan
    return {"seconds": dt, "ascans": ascans, "mhz": ascans / dt / 1e6, "threads": threads, "kind": kind,
            "sample": f"{bscans} B-scans ({ascans} A-scans) of the workload, reference CPU path "
                      f"(processor.tpp, FFTW-API substitute), {threads} thread(s)"}


def cpu_single_thread(q, bscans=4):
    """the reference's CPU path "as shipped" runs on one thread; a small sample is enough for a rate"""
    try:
        r = cpu_reference_run(q, 1, bscans, repeats=1, calibrate=False)
        return {"value": r["mhz"], "unit": "MHz (1e6 A-scans/s)", "sample": r["sample"]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--mode", default="fused", choices=["fused", "split", "cufft"])
    ap.add_argument("--enface", default="p2p", choices=["p2p", "nccl"], help="multi-GPU en-face gather: own peer-memory kernel or NCCL")
    ap.add_argument("--no-packed", action="store_true", help="skip the 12-bit packed-input extension measurement")
    ap.add_argument("--cpu-bscans", type=int, default=0, help="B-scans in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-numa", action="store_true", help="do not move the host thread to the GPU-local CPUs before pinning the host buffers")
    ap.add_argument("--separate-conversion", action="store_true",
                    help="end-to-end leg: floatToOutput as its own pass (the reference's order) instead of folded into the fused kernel's epilogue")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    n, a, b, bits = WORKLOADS[args.workload]
    from octproz_b200 import benchmark_params
    q = benchmark_params(n, a, b, bits)
    for k, v in WORKLOAD_PARAMS.get(args.workload, {}).items():
        setattr(q, k, v)
    q.update_all_curves()
    ncores = os.cpu_count() or 1
    ascans_per_step = a * b
    extra_chain = " + B-scan flip + sinusoidal scan correction" if q.sinusoidalScanCorrection else ""
    config = {"workload": f"{args.workload} volume, benchmark INI settings (cubic k-lin + dispersion + Hann + FPN once + log){extra_chain}, "
                          f"u16 container, synthetic", "samples_per_ascan": n, "ascans_per_bscan": a, "bscans_per_buffer": b,
              "bit_depth": bits}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        bscans = args.cpu_bscans or max(8, min(b, 2 * ncores))
        for _ in range(min(args.warmup, 1)):
            cpu_reference_run(q, ncores, bscans)
        t = []
        for _ in range(args.steps):
            t.append(cpu_reference_run(q, ncores, bscans))
        sec = float(np.mean([x["seconds"] for x in t])); mhz = t[0]["ascans"] / sec / 1e6
        line = {"impl": "reference", "metric": "A-scan rate (raw -> B-scan hot path)", "value": mhz, "unit": "MHz (1e6 A-scans/s)",
                "volumes_per_s": mhz * 1e6 / ascans_per_step, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config, step=f"bounded sample: {t[0]['ascans']} A-scans per step"),
                "cpu_baseline": {"value": mhz, "unit": "MHz (1e6 A-scans/s)", "cores": t[0]["threads"], "kind": t[0]["kind"], "sample": t[0]["sample"],
                                 "single_thread": cpu_single_thread(q)},
                "e2e": {"value": mhz, "unit": "MHz (1e6 A-scans/s)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "host_cores": ncores}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    from octproz_b200 import OctPipeline, _lib
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

    raw_np = [make_raw(q, seed_offset=16 * rank), make_raw(q, seed_offset=16 * rank + 8)]
    h_raw = [torch.from_numpy(x).pin_memory() for x in raw_np]
    d_raw = [x.cuda(non_blocking=False) for x in h_raw]             # two distinct 256 MiB inputs: larger than L2 (126 MB)
    bytes_in = raw_np[0].nbytes
    conv_bytes = (n // 2) * a * b * 2
    h_stream = [np.zeros(conv_bytes, np.uint8) for _ in range(2)]   # plain host memory; the library pins it like the reference (cuda_code.cu:661)

    qq = copy.deepcopy(q)
    p = OctPipeline(fft_mode=mode, device=local, bscan_index_base=(rank * b) % 2,
                    flags=_lib.FLAG_SEPARATE_CONVERSION if args.separate_conversion else 0)
    if not p.initializeCuda(None, None, qq):
        raise SystemExit("initializeCuda failed: " + getattr(p, "_create_error", ""))
    enface = torch.empty(a * b, dtype=torch.float32, device="cuda")
    gathered = torch.empty(world * a * b, dtype=torch.float32, device="cuda") if world > 1 else None
    stream = torch.cuda.ExternalStream(int(p._lib.octb200_compute_stream(p.handle)), device=torch.device("cuda", local))

    def sync_all():
        p.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def fpn_share():
        # FPN line of the first buffer: rank 0 determines, everyone uses it (8 KB broadcast)
        if dist is None:
            return
        ml = torch.from_numpy(p.fpn_mean_line()).cuda()
        dist.broadcast(ml, 0)
        p.set_fpn_mean_line(ml.cpu().numpy())

    # en-face slice of the G-times larger volume, every step: the library's own kernel stores each rank's slab straight into every
    # rank's frame window over NVLink peer memory (octb200_enface_gather); extraction + ncclAllGather only if IPC is unavailable
    gather_impl = None
    if dist is not None and args.enface != "nccl":
        try:
            mine = p.enface_gather_init(rank, world, world * a * b, rank * a * b)
            t = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            p.enface_gather_connect(b"".join(x.cpu().numpy().tobytes() for x in parts))
            ok = torch.ones(1, device="cuda")
        except Exception as e:  # noqa: BLE001
            print(f"rank {rank}: peer-memory en-face gather unavailable ({e}); using NCCL", file=sys.stderr, flush=True)
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_impl = "p2p" if float(ok.item()) > 0 else "nccl"
        dist.barrier()
    elif dist is not None:
        gather_impl = "nccl"

    if gather_impl == "p2p":
        p.enface_gather_auto(True, 100, 1, 0)      # every process call gathers depth 100: fused into the main kernel's epilogue

    def enface_step():
        if gather_impl == "p2p":
            pass
        elif gather_impl == "nccl":
            p.changeDisplayedEnFaceFrame(100, 1, 0, enface)
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, enface)

    def enface_finish():
        if gather_impl == "p2p":
            p.enface_gather_wait()         # every rank's slab of the last frame has arrived (flags, system-scope acquire)

    def step_resident(i):
        p.process_device(d_raw[i & 1])
        enface_step()

    # warm-up (includes LUT build, FPN determination, cuFFT plan if any)
    p.process_device(d_raw[0]); p.sync(); fpn_share()
    for i in range(max(3, args.warmup)):
        step_resident(i)
    sync_all()

    # ---- device-resident timed region: CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local); sampler.start()     # samples run until the end of the end-to-end region
    launches0 = p.launch_count()
    sync_all()
    p.event_record(0)
    for i in range(args.steps):
        step_resident(i)
    enface_finish()
    p.event_record(1)
    ms_total = p.event_elapsed_ms(0, 1)
    sync_all()
    launches = p.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * ascans_per_step / (ms_step * 1e3)   # MHz

    # ---- dominant kernel alone (roofline): events around back-to-back launches of the fused kernel ----
    kern_ms = p.time_kernel(d_raw[1], 20)
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

    # ---- end to end through octCudaPipeline(host buffer): pinned H2D + converted-output D2H inside the timed region ----
    p.sync()
    qq.streamToHost = True
    p.cuda_registerStreamingBuffers(h_stream[0], h_stream[1], conv_bytes)
    for i in range(max(3, args.warmup)):
        p.octCudaPipeline(h_raw[i & 1].numpy())
    sync_all()
    e2e_steps = max(1, min(args.steps, 200))
    e2e_launches0 = p.launch_count()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        p.octCudaPipeline(h_raw[i & 1].numpy())
        enface_step()
    enface_finish()
    p.sync(); torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    clocks = sampler.stop()
    e2e_mhz = world * ascans_per_step * e2e_steps / e2e_s / 1e6
    checksum = int(h_stream[0][:4096].view(np.uint16).sum())
    e2e_launches = (p.launch_count() - e2e_launches0) / e2e_steps
    p.cuda_unregisterStreamingBuffers()

    # ---- what the host link of this box can do (context for e2e): pinned H2D of one raw buffer alone, and with a D2H of the
    #      converted-output size running the other way at the same time (the steady state of the end-to-end loop) ----
    link = None
    try:
        d_probe = torch.empty_like(d_raw[0]); d_conv = torch.empty(conv_bytes, dtype=torch.uint8, device="cuda")
        h_conv = torch.empty(conv_bytes, dtype=torch.uint8).pin_memory()
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        def h2d_rate(with_d2h, reps=8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s_up):
                e0.record()
                for i in range(reps):
                    d_probe.copy_(h_raw[i & 1], non_blocking=True)
                e1.record()
            if with_d2h:
                with torch.cuda.stream(s_dn):
                    for i in range(2 * reps):
                        h_conv.copy_(d_conv, non_blocking=True)
            torch.cuda.synchronize()
            return bytes_in * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        h2d_rate(False, 2)
        alone, duplex = h2d_rate(False), h2d_rate(True)
        link = {"h2d_gbs_alone": alone, "h2d_gbs_with_concurrent_d2h": duplex,
                "e2e_h2d_gbs": bytes_in * e2e_steps / e2e_s / 1e9, "e2e_frac_of_duplex_h2d": bytes_in * e2e_steps / e2e_s / 1e9 / duplex}
        del d_probe, d_conv, h_conv
    except Exception as e:  # noqa: BLE001
        link = {"error": repr(e)}

    # ---- extension beside the headline (N = 1, 12-bit workload): the same volume delivered 12-bit PACKED (3 bytes per 2 samples,
    #      include/octb200.h OCTB200_PACK_12P; the reference only takes containers).  Reported separately, never as `value` / `e2e`. ----
    packed = None
    if world == 1 and bits == 12 and args.mode == "fused" and not args.no_packed:
        from octproz_b200.packing import pack12
        pp = OctPipeline(fft_mode=mode, device=local, input_packing=_lib.PACK_12P)
        qp = copy.deepcopy(q)
        if pp.initializeCuda(None, None, qp):
            hp = [torch.from_numpy(pack12(x)).pin_memory() for x in raw_np]
            dp = [x.cuda() for x in hp]
            pp.process_device(dp[0]); pp.sync()
            for i in range(5):
                pp.process_device(dp[i & 1])
            pp.sync()
            n_res = min(args.steps, 200)
            pp.event_record(0)
            for i in range(n_res):
                pp.process_device(dp[i & 1])
            pp.event_record(1)
            ms_p = pp.event_elapsed_ms(0, 1) / n_res
            qp.streamToHost = True
            hs2 = [np.zeros(conv_bytes, np.uint8) for _ in range(2)]
            pp.cuda_registerStreamingBuffers(hs2[0], hs2[1], conv_bytes)
            for i in range(3):
                pp.octCudaPipeline(hp[i & 1].numpy())
            pp.sync()
            n_e2e = min(args.steps, 100)
            t0 = time.perf_counter()
            for i in range(n_e2e):
                pp.octCudaPipeline(hp[i & 1].numpy())
            pp.sync()
            dt = time.perf_counter() - t0
            packed = {"input": "12-bit packed (Mono12p), extension", "value": ascans_per_step / (ms_p * 1e3), "ms_per_step": ms_p,
                      "e2e": {"value": ascans_per_step * n_e2e / dt / 1e6, "h2d_bytes_per_step": int(hp[0].numel()), "d2h_bytes_per_step": conv_bytes,
                              "ms_per_step": dt * 1e3 / n_e2e, "steps": n_e2e},
                      "unit": "MHz (1e6 A-scans/s)", "checksum": int(hs2[0][:4096].view(np.uint16).sum()) + int(hs2[1][:4096].view(np.uint16).sum())}
            pp.cuda_unregisterStreamingBuffers()
            pp.cleanupCuda()
            del hp, dp

    aff.__exit__(None, None, None)          # the CPU baseline gets every core again

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1:
        bscans = args.cpu_bscans or b            # one full volume per repeat: ~3 s of CPU work each
        c = cpu_reference_run(q, ncores, bscans, repeats=4)
        cpu = {"value": c["mhz"], "unit": "MHz (1e6 A-scans/s)", "cores": c["threads"], "kind": c["kind"], "sample": c["sample"],
               "single_thread": cpu_single_thread(q)}

    if dist is not None:
        p.sync(); torch.cuda.synchronize(); dist.barrier()       # peers have stopped writing into this rank's window
        if gather_impl == "p2p":
            p.enface_gather_close()
    p.cleanupCuda()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": "A-scan rate (raw -> B-scan hot path)", "value": value, "unit": "MHz (1e6 A-scans/s)",
            "volumes_per_s": value * 1e6 / ascans_per_step, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "ours",
            "config": dict(config, mode=args.mode, l2=f"two alternating {bytes_in >> 20} MiB inputs and {(n // 2) * a * b * 4 >> 20} MiB outputs per GPU: larger than the 126 MB L2",
                           parallelism=f"b-scan sharding x{world}" + ((", en-face slice gathered every step inside the fused kernel's epilogue over NVLink peer memory"
                                                                       if gather_impl == "p2p" else ", NCCL all-gather of the en-face slice every step") if world > 1 else "")),
            "e2e": {"value": e2e_mhz, "unit": "MHz (1e6 A-scans/s)", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": conv_bytes,
                    "ms_per_step": e2e_s * 1e3 / e2e_steps, "steps": e2e_steps, "timer": "host wall clock between device synchronisations, max over ranks",
                    "checksum": checksum, "gpu_launches_per_step": e2e_launches,
                    "conversion": "floatToOutput as a separate pass" if args.separate_conversion or args.mode != "fused"
                    else "floatToOutput folded into the fused kernel's epilogue (u16 line written beside the float line)",
                    "host_numa": host_numa, "link": link},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "host_cores": ncores}
    if packed is not None:
        line["packed12"] = packed
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
