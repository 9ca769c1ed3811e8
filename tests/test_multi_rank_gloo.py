"""N>1 path on the CPU: world_size 2, gloo.  Each rank renders its shard of a small sweep with the real
engine (kernel bodies through the block emulator), rank 0 gathers the buffers the way bench.py does with
NCCL, and the gathered audio is checked against the oracle render by render."""
import os
import socket
import sys

import numpy as np
import pytest

from audio_suite_b200 import parallel


def test_partition_covers_everything_once():
    for n, w in ((4096, 8), (4096, 3), (5, 8), (0, 2), (7, 1)):
        seen = []
        for r in range(w):
            seen += list(parallel.partition(n, w, r))
        assert seen == list(range(n))
        sizes = [len(parallel.partition(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_balanced_partition_is_a_partition():
    costs = np.random.default_rng(0).uniform(1, 10, 101)
    owners = parallel.balanced_partition(costs, 8)
    assert sorted(sum(owners, [])) == list(range(101))
    loads = [sum(costs[i] for i in o) for o in owners]
    assert max(loads) - min(loads) < 10.0


def test_balanced_equal_partition_keeps_counts_equal():
    costs = np.random.default_rng(1).uniform(1, 10, 4096)
    owners = parallel.balanced_equal_partition(costs, 8)
    assert sorted(sum(owners, [])) == list(range(4096)) and {len(o) for o in owners} == {512}
    loads = [sum(costs[i] for i in o) for o in owners]
    assert (max(loads) - min(loads)) / np.mean(loads) < 0.01
    from audio_suite_b200 import configs
    assert parallel.param_cost(configs.c5_params(3)) > 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sweep(i):
    from audio_suite_b200 import configs
    p = configs.c5_params(i)
    p["out_dur_s"] = 0.05 + 0.01 * (i % 3)            # ragged output lengths across ranks
    p["time_unfold"] = 25.0 + i
    p["micro_ms"] = 1.0
    return p


def _worker(rank, world, port, n_items, q):
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(here, "host_emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from emul_device import EmulDevice
    from audio_suite_b200 import engine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = list(parallel.partition(n_items, world, rank))
    br = engine.BatchRenderer([_sweep(i) for i in mine], device=EmulDevice())
    br.run()
    local = torch.from_numpy(np.asarray(br.out).copy())
    frames = [sum(int(round(_sweep(i)["out_dur_s"] * 48000)) for i in parallel.partition(n_items, world, r)) for r in range(world)]
    got = parallel.gather_frames(local, frames, dist, rank, world)
    # the overlapped form bench.py uses: equal slices per rank, gathered asynchronously one by one
    sg = parallel.SlicedGather([1000, 500], dist, rank, world, torch.device("cpu"))
    parts = [torch.full((2000,), float(rank + 1)), torch.full((1000,), float(10 * (rank + 1)))]
    for k, part in enumerate(parts):
        sg.start(k, part)
    recv = sg.finish()
    # the default at N>1: pass k is gathered while pass k+1 renders into the other output buffer
    pg = parallel.PipelinedGather(300, dist, rank, world, torch.device("cpu"))
    bufs = [torch.zeros(600), torch.zeros(600)]
    seen = []
    for k in range(5):
        s_ = pg.slot()
        if rank == 0 and k >= 2:
            seen.append([float(pg.recv[s_][r][0]) for r in range(world)])       # pass k-2, complete before its buffer is reused
        bufs[s_].fill_(float(100 * k + rank))
        pg.start(bufs[s_])
    recv2 = pg.finish()
    if rank == 0:
        assert [float(recv[0][r][0]) for r in range(world)] == [1.0, 2.0] and [float(recv[1][r][-1]) for r in range(world)] == [10.0, 20.0]
        assert seen == [[0.0, 1.0], [100.0, 101.0], [200.0, 201.0]]
        assert [float(recv2[0][r][0]) for r in range(world)] == [400.0, 401.0] and [float(recv2[1][r][0]) for r in range(world)] == [300.0, 301.0]
        q.put([g.numpy().copy() for g in got])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_matches_oracle():
    import torch.multiprocessing as mp
    from oracle import microsound_np as O
    n_items, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = np.concatenate(got)
    off = 0
    for i in range(n_items):
        ref, _ = O.render(_sweep(i))
        n = ref.shape[0]
        mine = flat[off:off + 2 * n].reshape(n, 2).astype(np.float64)
        assert np.max(np.abs(mine - ref)) < 1e-5, i
        off += 2 * n
    assert off == flat.size
