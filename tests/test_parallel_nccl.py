"""Multi-GPU numerical proof (SURVEY section 4 vii, 8e): N ranks through attach_data_parallel over NCCL hold, after the
all-reduce, exactly the mean of the gradients N independent single-GPU runs produce on the same shards
(per-replica BatchNorm statistics, like the reference's nn.DataParallel replicas, mainclean.py:132-134)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests.golden_util import load_case, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_nccl_gradient_equals_mean_of_shard_gradients(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "grads.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    got = [torch.load(out + ".rank%d" % k) for k in range(2)]
    assert torch.equal(got[0]["flat"], got[1]["flat"])            # every rank holds the same averaged gradient
    # the same two shards, one after the other, on one GPU without any process group
    from pathtracker_models_b200 import InT
    case = load_case("int_small")
    m = InT.InT(dimensions=32, timesteps=8, kernel_size=7)
    m.load_state_dict(case["sd"])
    m = m.cuda()
    x = torch.cat([case["x"].float(), case["x"].float().flip(0)], 0).cuda()
    y = torch.cat([case["y"].float(), case["y"].float().flip(0)], 0).cuda()
    B = x.shape[0] // 2
    flats = []
    for k in range(2):
        m.zero_grad(set_to_none=True)
        o, _ = m(x[k * B:(k + 1) * B])
        torch.nn.functional.binary_cross_entropy_with_logits(o, y[k * B:(k + 1) * B]).backward()
        flats.append(torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu())
        assert torch.equal(o.detach().cpu(), got[k]["logits"])    # same kernels, same inputs: bit-identical per rank
    mean = (flats[0].double() + flats[1].double()) / 2
    assert rel_err(got[0]["flat"], mean) < 1e-6
