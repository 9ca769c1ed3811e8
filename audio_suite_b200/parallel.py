"""Sharding a batch of independent renders over the GPUs of one box.

The reference's batch render is a serial loop of independent `render(p)` calls
(main_v2.py:1578-1593): no state is shared and nothing is exchanged, so the partition is by render
index, one process per GPU, identical kernel sequences, and the only collective is the final
collection of the rendered buffers on rank 0 (NCCL `gather`; `gloo` in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def partition(n_items: int, world: int, rank: int):
    """Contiguous block of render indices owned by `rank`: sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def balanced_partition(costs, world: int):
    """Greedy longest-processing-time assignment; returns a list of index lists, one per rank.
    Cost model of one render: grain FFT work n log n plus the per-output-frame tail."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    loads = np.zeros(world)
    owners = [[] for _ in range(world)]
    for i in order.tolist():
        r = int(np.argmin(loads))
        owners[r].append(i)
        loads[r] += costs[i]
    return [sorted(o) for o in owners]


def balanced_equal_partition(costs, world: int):
    """Equal COUNTS per rank (so the rendered slabs gather as equal-shape buffers), near-equal cost: renders sorted by
    cost are dealt to the ranks in snake order.  Returns a list of sorted index lists; the counts differ by at most one."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable").tolist()
    owners = [[] for _ in range(world)]
    for k, i in enumerate(order):
        rnd, pos = divmod(k, world)
        owners[pos if rnd % 2 == 0 else world - 1 - pos].append(i)
    return [sorted(o) for o in owners]


def param_cost(p):
    """Cost estimate of one render from its parameters alone (no planning): the per-output-frame tail plus, per expected
    event, the grain's FFT work.  Used to balance a sweep over the ranks before anything is planned."""
    base_sr = float(p["base_sr"])
    out_n = max(1.0, float(p["out_dur_s"]) * base_sr)
    n = max(16.0, base_sr * max(1.0, float(p["time_unfold"])) * float(p["micro_ms"]) / 1000.0)
    events = 1.0 if p["event_process"] == "Single" else max(1.0, min(float(p["max_grains"]), float(p["grains_per_sec"]) * float(p["out_dur_s"])))
    return 56.0 * out_n + events * n * (20.0 + 8.0 * float(np.log2(max(2.0, n))))


def render_cost(plan):
    c = 56.0 * plan.out_n
    for ev in plan.events:
        c += ev.n * (20.0 + 8.0 * np.log2(max(2, ev.n)))
    return c


def gather_frames(local, frames_per_rank, dist, rank, world, dst=0):
    """Collect per-rank interleaved stereo buffers (`local`: 1-D tensor of 2*frames floats) on rank
    `dst`.  Ranks may own different frame counts; buffers are padded to the largest.  Returns the list of
    per-rank tensors on `dst` (views trimmed to the true size) and None elsewhere."""
    import torch
    longest = 2 * max(frames_per_rank)
    if local.numel() < longest:
        pad = torch.zeros(longest, dtype=local.dtype, device=local.device)
        pad[:local.numel()] = local
        local = pad
    if rank == dst:
        bufs = [torch.empty(longest, dtype=local.dtype, device=local.device) for _ in range(world)]
        dist.gather(local, gather_list=bufs, dst=dst)
        return [b[:2 * f] for b, f in zip(bufs, frames_per_rank)]
    dist.gather(local, gather_list=None, dst=dst)
    return None


class SlicedGather:
    """Gather of per-rank rendered buffers that overlaps with rendering: every rank cuts its renders into `slices`
    equal parts; as soon as slice k has been rendered its buffer is gathered on rank `dst` (NCCL runs the collective
    on its own stream) while slice k+1 renders.  Receive buffers are allocated once."""

    def __init__(self, frames_per_slice, dist, rank, world, device, dst=0):
        import torch
        self.dist, self.rank, self.world, self.dst = dist, rank, world, dst
        self.frames = list(frames_per_slice)
        self.recv = None
        if rank == dst:
            self.recv = [[torch.empty(2 * f, dtype=torch.float32, device=device) for _ in range(world)] for f in self.frames]
        self.pending = []

    def start(self, k, local):
        """local: this rank's 1-D float32 buffer of slice k (2 * frames[k] values on every rank)."""
        if local.numel() != 2 * self.frames[k]:
            raise ValueError("slice buffers must have the same size on every rank")
        w = self.dist.gather(local, gather_list=self.recv[k] if self.rank == self.dst else None, dst=self.dst, async_op=True)
        self.pending.append(w)

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        return self.recv


class _EventWork:
    """wait(): the current stream waits for the recorded event (the interface of a c10d Work as PipelinedGather uses it)."""

    def __init__(self, ev):
        self.ev = ev

    def wait(self):
        self.ev.wait()


class PipelinedGather:
    """Gather of per-rank rendered buffers that overlaps with the NEXT pass over the batch: the ranks render pass k+1 into
    the other of two output buffers while NCCL collects pass k on rank `dst` (the collective runs on NCCL's own stream
    behind an event of the rendering stream).  Back-to-back sweeps -- the reference's batch loop run again with the next
    base preset -- then cost max(render, gather) per pass instead of their sum, without cutting the batch into slices
    (measured on B200: slices render less efficiently -- 20.3 ms unsliced vs 23.6 ms in four slices at N = 2)."""

    def __init__(self, frames, dist, rank, world, device, dst=0, depth=2, collective="gather"):
        """collective: "gather" (NCCL point-to-point receives on rank `dst`) or "all_gather" (every rank ends up with every
        slab -- more traffic in total, but it runs on NCCL's ring / NVLS all-gather path instead of seven concurrent
        point-to-point streams into one GPU)."""
        import torch
        self.dist, self.rank, self.world, self.dst, self.depth = dist, rank, world, dst, depth
        self.frames = int(frames)
        self.collective = collective
        self.recv = None
        if collective == "all_gather":
            self.full = [torch.empty(world * 2 * self.frames, dtype=torch.float32, device=device) for _ in range(depth)]
            self.recv = [list(f.view(world, 2 * self.frames).unbind(0)) for f in self.full]
        elif collective == "peer_copy":
            # Rank `dst`'s receive slabs live in symmetric memory that every rank maps; a rank stores its slab with ONE
            # device-to-device copy over NVLink (copy engines: no SMs taken from the rendering kernels, unlike NCCL's
            # send / receive kernels), bracketed by the symmetric-memory barrier on a side stream.
            import torch.distributed._symmetric_memory as symm
            n = depth * world * 2 * self.frames
            self._sym = symm.empty(n, dtype=torch.float32, device=device)
            self._hdl = symm.rendezvous(self._sym, dist.group.WORLD)
            self._remote = self._hdl.get_buffer(dst, (depth, world, 2 * self.frames), torch.float32)
            self._side = torch.cuda.Stream(device=device)
            self._torch = torch
            if rank == dst:
                self.recv = [list(self._sym.view(depth, world, 2 * self.frames)[d].unbind(0)) for d in range(depth)]
        elif rank == dst:
            self.recv = [[torch.empty(2 * self.frames, dtype=torch.float32, device=device) for _ in range(world)] for _ in range(depth)]
        self.work = [None] * depth
        self.k = 0

    def slot(self):
        """Index of the output buffer the next pass may render into (its previous gather has completed)."""
        s = self.k % self.depth
        if self.work[s] is not None:
            self.work[s].wait()
            self.work[s] = None
        return s

    def start(self, local):
        """local: this rank's 1-D float32 buffer of the pass just enqueued (2 * frames values on every rank)."""
        if local.numel() != 2 * self.frames:
            raise ValueError("buffers must have the same size on every rank")
        s = self.k % self.depth
        if self.collective == "all_gather":
            self.work[s] = self.dist.all_gather_into_tensor(self.full[s], local, async_op=True)
        elif self.collective == "peer_copy":
            torch = self._torch
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self._hdl.barrier(channel=0)          # rank dst has let go of slot s (its consumer ran before this call)
                self._remote[s, self.rank].copy_(local, non_blocking=True)
                self._hdl.barrier(channel=1)          # every rank's store has landed
                ev = torch.cuda.Event()
                ev.record()
            self.work[s] = _EventWork(ev)
        else:
            self.work[s] = self.dist.gather(local, gather_list=self.recv[s] if self.rank == self.dst else None, dst=self.dst, async_op=True)
        self.k += 1

    def finish(self):
        for s in range(self.depth):
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None
        return self.recv
