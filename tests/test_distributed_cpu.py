"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: tensor sharding of the merge, request sharding of
the prefill, the max-over-ranks timing reduction, and the reference arm under torchrun (rank 0 prints, others exit 0)."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from modelcompose_b200 import synthetic as syn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = syn.dense_7b_tensor_shapes()
    sizes = [int(torch.Size(s).numel()) for _, s in shapes]
    mine = syn.shard_tensors_greedy(sizes, world)[rank]
    # every tensor is owned by exactly one rank
    owner = torch.zeros(len(sizes), dtype=torch.int64)
    owner[mine] = 1
    dist.all_reduce(owner)
    assert bool((owner == 1).all())
    # per-rank algorithmic bytes add up to the whole job; imbalance stays small
    my_bytes = torch.tensor([sum(sizes[i] for i in mine) * 8.0], dtype=torch.float64)
    total = my_bytes.clone()
    dist.all_reduce(total)
    assert total.item() == sum(sizes) * 8.0
    mx = my_bytes.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    assert mx.item() <= total.item() / world * 1.02
    # timing reduction used by bench.py: the slowest rank defines the step time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    # prefill request sharding: request 0 is the cross-rank probe (identical), the rest are rank-specific
    probe = syn.make_prompt_ids(1, ["vision", "audio"], 88, 32000, 3, {"vision": -200, "audio": -203}, 36)
    own = syn.make_prompt_ids(3, ["vision", "audio"], 88, 32000, 30 + rank, {"vision": -200, "audio": -203}, 36)
    gathered = [torch.empty_like(probe) for _ in range(world)]
    dist.all_gather(gathered, probe)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    others = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(others, own)
    assert not torch.equal(others[0], others[1])
    assert probe.shape[1] == 36 + 2 * 3 + 88 == 130
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_sharding_and_reductions_world2(tmp_path):
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_greedy_shard_balance(world):
    sys.path.insert(0, ROOT)
    from modelcompose_b200 import synthetic as syn
    sizes = [int(torch.Size(s).numel()) for _, s in syn.dense_7b_tensor_shapes()]
    parts = syn.shard_tensors_greedy(sizes, world)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) <= sum(sizes) / world * 1.02


def test_reference_arm_under_torchrun_world2():
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1", "--workload", "merge"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "GB/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
