"""torchrun check: the three PipelinedGather collectives deliver the same slabs on rank 0 (run on >= 2 GPUs)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_suite_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
frames = 1 << 20
ok = True
for mode in ("gather", "all_gather", "peer_copy"):
    pg = parallel.PipelinedGather(frames, dist, rank, world, dev, collective=mode)
    for k in range(5):
        s = pg.slot()
        g = torch.Generator(device=dev).manual_seed(1000 * k + rank)
        x = torch.rand(2 * frames, generator=g, device=dev, dtype=torch.float32)
        pg.start(x)
    recv = pg.finish()
    torch.cuda.synchronize()
    if rank == 0:
        for k in (3, 4):
            for r in range(world):
                g = torch.Generator(device=dev).manual_seed(1000 * k + r)
                want = torch.rand(2 * frames, generator=g, device=dev, dtype=torch.float32)
                same = bool(torch.equal(recv[k % 2][r], want))
                ok &= same
                if not same:
                    print("MISMATCH", mode, k, r)
        print("mode", mode, "checked")
    dist.barrier()
if rank == 0:
    print("GATHER_MODES_OK" if ok else "GATHER_MODES_FAILED")
dist.destroy_process_group()
