"""CPU, world_size 2, gloo: the host-side data-parallel logic (flat-gradient averaging hook,
parameter broadcast, disjoint shard indices).  The CUDA time loop itself is covered by -m gpu tests."""
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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pathtracker_models_b200 import InT
    from pathtracker_models_b200.parallel import (attach_data_parallel, average_flat_gradient_,
                                                  broadcast_parameters, shard_first_index)
    torch.manual_seed(100 + rank)                      # different init per rank ...
    m = InT.InT(dimensions=32, timesteps=4, kernel_size=7)
    broadcast_parameters(m, src=0)                     # ... made identical by the broadcast
    flat = m.flat_parameters()
    ref = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(ref, flat.detach())
    same = all(torch.equal(ref[0], r) for r in ref)
    # the hook path: a stand-in "loss" whose flat gradient differs per rank
    attach_data_parallel(m)
    flat = m.flat_parameters()
    flat.register_hook(lambda g: average_flat_gradient_(g.contiguous(), m._dp_group))
    w = torch.arange(flat.numel(), dtype=torch.float32) * (rank + 1)
    (flat * w).sum().backward()
    g = torch.cat([p.grad.reshape(-1) for n, p in m.named_parameters() if n != "unit1.w"])
    expect = torch.arange(flat.numel(), dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    idx = [shard_first_index(s, rank, world, 64) for s in range(3)]
    q.put((rank, same, torch.allclose(g, expect), m.unit1.w.grad is None, idx))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    for rank, same, ok, dead_none, idx in res:
        assert same and ok and dead_none, (rank, same, ok, dead_none)
    all_idx = [i for r in res for i in r[4]]
    assert len(set(all_idx)) == len(all_idx)                       # disjoint shards
    assert all(b - a >= 64 for a, b in zip(sorted(all_idx), sorted(all_idx)[1:]))
