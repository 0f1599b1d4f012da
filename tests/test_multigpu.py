"""GPU (-m gpu), needs >= 2 visible devices (skipped otherwise; run with ``gpurun --gpus 2``): the multi-rank step over
NCCL -- CapturedStep's three-graph path with the two collectives between the replays, bit-identical targets on every rank
and equal to the oracle's simulated world (tests/stepcheck.py::check_multi_rank), and GradBucket.sync over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
import fairguide
from tests import stepcheck
for kind, dtype, S in (("gender_race_age", torch.bfloat16, 100), ("gender_race", torch.float32, 40), ("gender", torch.float32, 1)):
    rep = stepcheck.check_multi_rank(kind, dtype, dev, rank, world, n_side_global=256, S=S, captured=True)
    if rank == 0:
        print("multi-rank", kind, dtype, rep)
# bucketed gradient synchronisation over NCCL against the per-tensor loop of E1:1999-2011
torch.manual_seed(rank)
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in ((50, 320), (320,), (7, 3, 3), (1,))]
for p in params:
    p.grad = torch.randn_like(p)
ref = []
for p in params:
    g = p.grad.clone(); dist.all_reduce(g); ref.append(g / world / 3)
ok = fairguide.allreduce_average_gradients(params, num_processes=world, n_backward=3)
assert ok
for p, r in zip(params, ref):
    assert torch.allclose(p.grad, r, rtol=3e-7, atol=0), (p.grad - r).abs().max()
params[1].grad[5] = float("inf")
flags = [fairguide.allreduce_average_gradients(params, num_processes=world, n_backward=3)]
assert flags[0] is False
dist.barrier()
dist.destroy_process_group()
print("worker", rank, "ok")
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_nccl_ranks(tmp_path):
    port = 29500 + os.getpid() % 400
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out[-4000:]
        assert "ok" in out
