"""worker of tests/test_parallel_nccl.py (launched under torch.distributed.run, one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden_util import load_case  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pathtracker_models_b200 import InT
    from pathtracker_models_b200.parallel import attach_data_parallel, broadcast_parameters
    case = load_case("int_small")
    m = InT.InT(dimensions=32, timesteps=8, kernel_size=7)
    if rank == 0:
        m.load_state_dict(case["sd"])          # the other ranks start from their own random init: broadcast must fix that
    m = m.cuda()
    broadcast_parameters(m)
    attach_data_parallel(m)
    # global batch = the fixture's clips followed by the same clips flipped; rank r owns samples [r*B, (r+1)*B)
    x = torch.cat([case["x"].float(), case["x"].float().flip(0)], 0)
    y = torch.cat([case["y"].float(), case["y"].float().flip(0)], 0)
    B = x.shape[0] // world
    xs, ys = x[rank * B:(rank + 1) * B].cuda(), y[rank * B:(rank + 1) * B].cuda()
    out, _ = m(xs)
    torch.nn.functional.binary_cross_entropy_with_logits(out, ys).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu()
    torch.save({"flat": flat, "logits": out.detach().cpu()}, out_path + ".rank%d" % rank)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
