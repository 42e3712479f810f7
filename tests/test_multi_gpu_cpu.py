"""N > 1 host logic on CPU: world_size-2 gloo run of the prompt sharding + the path's single collective
(all-gather of finished waveforms).  The data path has no other exchange."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_prompts, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ma3_b200.pipeline import gather_waveforms, shard_prompts
    mine = shard_prompts(n_prompts, rank, world)
    # stand-in "waveforms": prompt index encoded in the samples
    wav = torch.stack([torch.full((16,), float(i)) for i in mine])
    allw = gather_waveforms(wav)
    if rank == 0:
        torch.save(allw, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    out = str(tmp_path / "gathered.pt")
    n_prompts, world = 8, 2
    mp.spawn(_worker, args=(world, _free_port(), n_prompts, out), nprocs=world, join=True)
    allw = torch.load(out)
    assert allw.shape == (n_prompts, 16)
    # rank-major order: rank 0's prompts (0,2,4,6) then rank 1's (1,3,5,7); every prompt exactly once
    assert allw[:, 0].tolist() == [0, 2, 4, 6, 1, 3, 5, 7]


def test_gather_single_process_is_identity():
    from ma3_b200.pipeline import gather_waveforms
    w = torch.randn(3, 5)
    assert gather_waveforms(w) is w
