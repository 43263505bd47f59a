"""Worker of tests/test_gpu_multi.py, one process per GPU (torchrun): shard a small volume by B-scan, process it, gather the
en-face frame (a) with extraction + NCCL all_gather and (b) with the library's own peer-memory kernel, compare both with the
un-sharded oracle result, and time the two gathers.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octproz_b200 import benchmark_params, synth  # noqa: E402
from octproz_b200.sharding import ShardedPipeline  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (checker only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, a, btot = 1024, 64, 8 * world
    q = benchmark_params(n, a, btot)
    q.bscanFlip = True
    q.fixedPatternNoiseRemoval = False
    q.update_all_curves()
    raw = synth.make_volume(n, a, btot, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, _, _ = orc.process(q, raw)
    sp = ShardedPipeline(q, rank, world, dist=dist, device=local)
    sp._coll_device = dev
    assert sp.initialize()
    sp.process_host(sp.local_slice(raw)); sp.sync()
    out = {"world": world}
    worst = 0.0
    for (frame, nf, fn) in ((17, 1, 0), (100, 5, 0), (300, 4, 1)):
        ref_enface = orc.enface_frame(ref, n // 2, a, btot, frame, nf, fn)
        # (a) extraction kernel + NCCL all_gather
        def extract(f, k, m):
            t = torch.empty(a * sp.count, dtype=torch.float32, device=dev)
            sp.pipe.changeDisplayedEnFaceFrame(f, k, m, t); sp.pipe.sync()
            return t
        full = sp.enface(frame, nf, fn, extract).cpu().numpy()
        # (b) one kernel over peer memory
        if not getattr(sp, "_p2p", False):
            sp.connect_enface_peers(device=dev)
        ptr = sp.enface_p2p(frame, nf, fn); sp.sync()
        got = torch.empty(a * btot, dtype=torch.float32, device=dev)
        # wrap the library's frame window as a tensor through the CUDA array interface
        class _W:  # noqa: N801
            __cuda_array_interface__ = {"shape": (a * btot,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        got.copy_(torch.as_tensor(_W(), device=dev))
        got = got.cpu().numpy()
        e_nccl = float(np.abs(full - ref_enface).max()); e_p2p = float(np.abs(got - ref_enface).max())
        same = bool(np.array_equal(full, got))
        worst = max(worst, e_nccl, e_p2p)
        assert same, f"rank {rank}: peer gather differs from all_gather for frame {(frame, nf, fn)}"
        assert e_p2p < 1e-3, f"rank {rank}: en-face differs from the oracle by {e_p2p}"
        dist.barrier()
    # ---- automatic gather inside the fused kernel's epilogue: process + gather in ONE launch, every rank's frame complete ----
    sp.pipe.enface_gather_auto(True, 17, 1, 0)
    sp.pipe.octCudaPipeline(sp.local_slice(raw))
    ptr = sp.pipe.enface_gather_wait(); sp.sync()

    class _W2:  # noqa: N801
        __cuda_array_interface__ = {"shape": (a * btot,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    got = torch.as_tensor(_W2(), device=dev).clone().cpu().numpy()
    e_auto = float(np.abs(got - orc.enface_frame(ref, n // 2, a, btot, 17, 1, 0)).max())
    assert e_auto < 1e-3, f"rank {rank}: fused en-face gather differs from the oracle by {e_auto}"
    out["fused_gather_max_abs_err_vs_oracle"] = e_auto
    sp.pipe.enface_gather_auto(False)
    dist.barrier()
    # ---- flow control: rank 0 runs ahead (its host enqueues 12 gathers of DIFFERENT depth frames back to back), the last rank dawdles
    #      between its gathers; every frame every rank hands out must still be exactly the frame of that sequence number ----
    stream = torch.cuda.ExternalStream(int(sp.pipe._lib.octb200_compute_stream(sp.pipe.handle)), device=dev)
    import time
    frames = list(range(20, 32))
    snaps = []
    for f in frames:
        ptr = sp.enface_p2p(f, 1, 0)

        class _W3:  # noqa: N801
            __cuda_array_interface__ = {"shape": (a * btot,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        with torch.cuda.stream(stream):                  # stream ordered behind the consumer kernel of this gather
            snaps.append(torch.as_tensor(_W3(), device=dev).clone())
        if rank == world - 1:
            sp.sync(); time.sleep(0.02)
    sp.sync(); torch.cuda.synchronize()
    for f, snap in zip(frames, snaps):
        want = orc.enface_frame(ref, n // 2, a, btot, f, 1, 0)
        e = float(np.abs(snap.cpu().numpy() - want).max())
        assert e < 1e-3, f"rank {rank}: frame {f} torn or stale under skew (max err {e})"
    st = sp.pipe.enface_gather_status()
    assert st["ack_timeouts"] == 0 and st["arrival_timeouts"] == 0, st
    out["flow_control_frames_checked"] = len(frames)
    dist.barrier()
    # ---- back-to-back SHORT buffers with the gather fused into every kernel (the strong-scaling regime: a few lines per SM and buffer):
    #      the producers' flow control must never wait for a consumer kernel that cannot get an SM (status counters stay 0, no stall) ----
    sp.pipe.enface_gather_auto(True, 17, 1, 0)
    d_local = torch.from_numpy(np.ascontiguousarray(sp.local_slice(raw)).view(np.int16)).to(dev)
    sp.pipe.process_device(d_local); sp.sync(); dist.barrier()
    t0 = time.perf_counter()
    raw_call, h, ptr_local = sp.pipe._lib.octb200_process_device, sp.pipe.handle, int(d_local.data_ptr())
    for _ in range(2000):               # the bare C call: the host stays far ahead of the GPU, the compute stream is saturated
        assert raw_call(h, ptr_local) == 0
    ptr = sp.pipe.enface_gather_wait(); sp.sync()
    dt = time.perf_counter() - t0
    st = sp.pipe.enface_gather_status()
    assert st["ack_timeouts"] == 0 and st["arrival_timeouts"] == 0, f"rank {rank}: {st}"
    assert dt < 5.0, f"rank {rank}: 2000 short buffers took {dt:.1f} s"

    class _W4:  # noqa: N801
        __cuda_array_interface__ = {"shape": (a * btot,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    got = torch.as_tensor(_W4(), device=dev).clone().cpu().numpy()
    assert float(np.abs(got - orc.enface_frame(ref, n // 2, a, btot, 17, 1, 0)).max()) < 1e-3
    out["short_buffers_us_per_step"] = dt / 2000 * 1e6
    sp.pipe.enface_gather_auto(False)
    dist.barrier()
    # ---- timing of the two gathers (device events, max over ranks) ----
    loc = torch.empty(a * sp.count, dtype=torch.float32, device=dev)
    gathered = torch.empty(world * a * sp.count, dtype=torch.float32, device=dev)
    iters = 50
    def t_nccl():
        sp.pipe.changeDisplayedEnFaceFrame(17, 1, 0, loc)
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(gathered, loc)
    def t_p2p():
        sp.pipe.enface_gather(17, 1, 0)
    for name, fn_ in (("nccl_us", t_nccl), ("p2p_us", t_p2p)):
        for _ in range(5):
            fn_()
        sp.sync(); dist.barrier(); torch.cuda.synchronize()
        sp.pipe.event_record(0)
        for _ in range(iters):
            fn_()
        sp.pipe.event_record(1)
        ms = sp.pipe.event_elapsed_ms(0, 1)
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t.item()) * 1e3 / iters
        sp.sync(); dist.barrier()
    out["max_abs_err_vs_oracle"] = worst
    sp.close_enface_peers()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
