"""Multi-GPU host side: shard a buffer by B-scan, one process per GPU.

Every stage of the path is per A-scan or per B-scan, so the shards never exchange sample data.  What crosses ranks:
  * the fixed-pattern-noise line (N complex floats, 8 KB): determined by rank 0 from the first B-scans of the buffer
    (cuda_code.cu:1520-1522) and broadcast once (or per buffer in continuous mode);
  * the displayed en-face slice: each rank extracts its A x B_local floats, one all_gather assembles the frame
    (the only collective on the display path; the volume itself stays sharded in HBM);
  * the 3-D volume view, when enabled: each rank's u8 voxels, one all_gather (`volume_view`).
The en-face gather has two implementations: `enface` (local extraction + one all_gather: NCCL / gloo) and
`connect_enface_peers` + `enface_p2p` (the library's own kernel stores every value straight into all ranks' frame windows
over NVLink peer memory and publishes a sequence flag: one kernel, no NCCL call on the display path).
`dist` is a torch.distributed-like module (NCCL on GPUs, gloo in the CPU tests); `pipeline_factory` builds the
per-rank engine (OctPipeline on a B200; the tests inject an oracle-backed stand-in to check the host logic).
"""
from __future__ import annotations

import copy

import numpy as np


def shard_bounds(total_bscans: int, world: int, rank: int) -> tuple[int, int]:
    """B-scans [start, start+count) of rank `rank`.  Shares are equal when world divides total; starts are kept even
    where possible so that the parity of 'flip every even B-scan' (cuda_code.cu:795) needs no special casing."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total_bscans, world)
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    start = sum(counts[:rank])
    return start, counts[rank]


class ShardedPipeline:
    def __init__(self, params, rank: int, world: int, dist=None, pipeline_factory=None, device: int = -1, fft_mode: int = 0):
        self.rank, self.world, self.dist = rank, world, dist
        self.full = params
        self.start, self.count = shard_bounds(int(params.bscansPerBuffer), world, rank)
        self.local = copy.deepcopy(params)
        self.local.bscansPerBuffer = self.count
        if pipeline_factory is None:
            from .pipeline import OctPipeline
            total = int(params.bscansPerBuffer)
            pipeline_factory = lambda base: OctPipeline(fft_mode=fft_mode, device=device, bscan_index_base=base,  # noqa: E731
                                                        bscans_in_unsharded_buffer=total)
        # the index of the shard's first B-scan in the un-sharded buffer: flip parity AND the reference's "last B-scan of an odd buffer
        # is never flipped" (cuda_code.cu:794-805) are decided on un-sharded indices
        self.pipe = pipeline_factory(self.start)
        self._fpn_shared = False

    def initialize(self, h1=None, h2=None) -> bool:
        if self.count == 0:
            return True
        return self.pipe.initializeCuda(h1, h2, self.local)

    def local_slice(self, full_raw: np.ndarray) -> np.ndarray:
        q = self.full
        v = full_raw.reshape(q.bscansPerBuffer, q.ascansPerBscan, q.samplesPerLine)
        return np.ascontiguousarray(v[self.start:self.start + self.count])

    # ---- FPN line: rank 0 determines, everybody uses it ----
    def _broadcast_fpn(self):
        import torch
        n = int(self.full.samplesPerLine)
        if self.rank == 0:
            ml = torch.from_numpy(np.ascontiguousarray(self.pipe.fpn_mean_line(), np.float32))
        else:
            ml = torch.zeros((n, 2), dtype=torch.float32)
        dev = getattr(self, "_coll_device", None)
        if dev is not None:
            ml = ml.to(dev)
        self.dist.broadcast(ml, 0)
        if self.rank != 0 and self.count:
            self.pipe.set_fpn_mean_line(ml.cpu().numpy())

    def process_host(self, h_raw_local) -> None:
        q = self.local
        fpn = bool(q.fixedPatternNoiseRemoval) and self.world > 1 and self.dist is not None
        need_share = fpn and (not self._fpn_shared or q.continuousFixedPatternNoiseDetermination or q.redetermineFixedPatternNoise)
        if fpn:
            _, count0 = shard_bounds(int(self.full.bscansPerBuffer), self.world, 0)
            if int(q.bscansForNoiseDetermination) > count0:
                raise ValueError(f"bscansForNoiseDetermination = {q.bscansForNoiseDetermination} exceeds rank 0's shard ({count0} B-scans): the "
                                 "fixed-pattern-noise line is determined from the FIRST B-scans of the buffer (cuda_code.cu:1520-1522)")
        if need_share:
            # rank 0 owns the first B-scans of the buffer (cuda_code.cu:1520): it runs first, then the line is shared
            if self.rank == 0:
                self.pipe.octCudaPipeline(h_raw_local); self.pipe.sync()
            self._broadcast_fpn()
            self._fpn_shared = True
            if self.rank != 0 and self.count:
                # the other ranks USE the broadcast line: with the determination flags still set their pipeline would re-determine
                # it from their own shard (run_chain, cuda_code.cu:1521) and overwrite it
                cont, q.continuousFixedPatternNoiseDetermination = q.continuousFixedPatternNoiseDetermination, False
                q.redetermineFixedPatternNoise = False
                try:
                    self.pipe.octCudaPipeline(h_raw_local)
                finally:
                    q.continuousFixedPatternNoiseDetermination = cont
            q.redetermineFixedPatternNoise = False
        elif self.count:
            self.pipe.octCudaPipeline(h_raw_local)

    def sync(self) -> None:
        if self.count:
            self.pipe.sync()

    # ---- en-face frame of the whole (sharded) volume ----
    def enface(self, frame_nr: int, n_frames: int, fn: int, local_extract, device=None):
        """local_extract(frame_nr, n_frames, fn) -> torch tensor [A*B_local] in the reference's order
        (disp[(E-1)-i], cuda_code.cu:909).  Returns the full frame [A*B_total] on every rank."""
        import torch
        a, btot = int(self.full.ascansPerBscan), int(self.full.bscansPerBuffer)
        mine = local_extract(frame_nr, n_frames, fn)
        if self.world == 1 or self.dist is None:
            return mine
        maxc = max(shard_bounds(btot, self.world, r)[1] for r in range(self.world))
        pad = torch.zeros(a * maxc, dtype=torch.float32, device=mine.device)
        pad[: mine.numel()] = mine
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad)
        # the reference writes the frame reversed (index (E-1)-i): global order = shards in reverse rank order
        out = [parts[r][: a * shard_bounds(btot, self.world, r)[1]] for r in reversed(range(self.world))]
        return torch.cat(out)

    # ---- 3-D volume view of the whole (sharded) volume ----
    def volume_view(self, local_tex):
        """local_tex: torch uint8 [N/2][B_local][A], this rank's voxels as octb200_volume_u8 lays them out (the GL_R8 3-D texture of
        updateDisplayedVolume, cuda_code.cu:915-941: x = A-scan, y = B-scan, z = flipped depth).  Returns the texture of the whole
        volume [N/2][B_total][A] on every rank: one all_gather (64 MiB at 1024 x 512 x 256); the shards are slabs
        in y, so the gathered parts are concatenated along the B-scan axis of every depth plane."""
        import torch
        a, btot = int(self.full.ascansPerBscan), int(self.full.bscansPerBuffer)
        h = int(self.full.samplesPerLine) // 2
        mine = local_tex.reshape(h, self.count, a)
        if self.world == 1 or self.dist is None:
            return mine
        counts = [shard_bounds(btot, self.world, r)[1] for r in range(self.world)]
        pad = torch.zeros((h, max(counts), a), dtype=torch.uint8, device=mine.device)
        pad[:, : self.count] = mine
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad)
        return torch.cat([parts[r][:, : counts[r]] for r in range(self.world)], dim=1)

    # ---- en-face frame gathered by the library's own kernel over peer memory (include/octb200.h: octb200_enface_gather_*) ----
    def connect_enface_peers(self, device=None) -> None:
        """collective: every rank allocates its frame window, the 64-byte IPC handles are exchanged rank-major with one
        all_gather, every rank opens the windows of its peers.  `device`: where the handle tensor lives for the collective
        (a CUDA device for NCCL, None/cpu for gloo)."""
        import torch
        if self.count == 0:
            raise ValueError("a rank without B-scans cannot take part in the peer gather")
        a, btot = int(self.full.ascansPerBscan), int(self.full.bscansPerBuffer)
        mine = self.pipe.enface_gather_init(self.rank, self.world, a * btot, a * self.start)
        if len(mine) != 64:
            raise ValueError("IPC handle must be 64 bytes")
        if self.world == 1 or self.dist is None:
            handles = mine
        else:
            t = torch.tensor(list(mine), dtype=torch.uint8)
            if device is not None:
                t = t.to(device)
            parts = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(parts, t)
            handles = b"".join(bytes(x.cpu().numpy().tobytes()) for x in parts)
        self.pipe.enface_gather_connect(handles)
        self._p2p = True
        if self.world > 1 and self.dist is not None:
            self.dist.barrier()           # nobody gathers into a window that is not open yet

    def enface_p2p(self, frame_nr: int, n_frames: int, fn: int, wait: bool = True) -> int:
        """extraction + peer stores + flag (one kernel on the compute stream).  With wait=True also enqueues the wait for every
        rank's slab and returns the device address of the assembled frame [A*B_total] floats in the reference's order."""
        if not getattr(self, "_p2p", False):
            raise RuntimeError("connect_enface_peers() first")
        self.pipe.enface_gather(frame_nr, n_frames, fn)
        return self.pipe.enface_gather_wait() if wait else 0

    def close_enface_peers(self) -> None:
        if getattr(self, "_p2p", False):
            self.sync()
            if self.world > 1 and self.dist is not None:
                self.dist.barrier()       # peers have stopped writing into this rank's window
            self.pipe.enface_gather_close()
            self._p2p = False
