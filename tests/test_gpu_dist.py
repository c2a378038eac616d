"""NCCL (not gloo) test of config 5's exchange step on real GPUs: retrieval.gather_rows with UNEQUAL contiguous shards, and the
per-step all-gather of padded detections.  Needs >= 2 GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`);
skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["WD_ROOT"])
import torch, torch.distributed as dist
from wedetect_b200 import dist as wdist
from wedetect_b200.retrieval import gather_rows
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True
for total, K in ((7, 1203), (1, 16), (64, 80), (5, 8)):        # 7 -> shards of 4 / 3; 1 -> 1 / 0 (an empty shard)
    mine = wdist.shard_indices(total, world, rank)
    local = torch.stack([torch.arange(K, dtype=torch.float32) + 1000.0 * i for i in mine]).to(dev) if len(mine) else torch.zeros(0, K, device=dev)
    full = gather_rows(local, total)
    want = torch.stack([torch.arange(K, dtype=torch.float32) + 1000.0 * i for i in range(total)]).to(dev)
    ok = ok and full.shape == want.shape and bool(torch.equal(full, want))
# the per-step exchange of the detection bench: one all-gather of the padded [B, M+1, 6] block
B, M = 4, 300
g = torch.Generator().manual_seed(7)
allb, alls = torch.rand(world * B, M, 4, generator=g), torch.rand(world * B, M, generator=g)
alll, allc = torch.randint(0, 80, (world * B, M), generator=g), torch.randint(0, M + 1, (world * B,), generator=g)
sl = slice(rank * B, (rank + 1) * B)
out = wdist.gather_detections(wdist.pack_detections(allb[sl].to(dev), alls[sl].to(dev), alll[sl].to(dev), allc[sl].to(dev)))
b, s, l, c = wdist.unpack_detections(out.cpu())
ok = ok and torch.equal(b, allb) and torch.equal(s, alls) and torch.equal(l, alll) and torch.equal(c, allc)
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL_GATHER_OK" if float(res) == 1.0 else "NCCL_GATHER_MISMATCH", dist.get_backend(), world)
dist.destroy_process_group()
'''


def test_gather_rows_unequal_shards_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, WD_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(w)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "NCCL_GATHER_OK nccl 2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
