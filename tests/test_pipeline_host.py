"""CPU: data-parallel host logic (sharding + host gather) on a 2-process gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocr_vi_invoice_b200.pipeline import gather_results, shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 7, 16, 20000):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = []
        for i in shard_indices(n_items, rank, world):
            boxes = [np.full((4, 2), i * 10 + k, dtype=np.int64) for k in range(i % 3)]  # ragged, some empty
            local.append((i, (boxes, [0.5 + 0.01 * k for k in range(i % 3)])))
        merged = gather_results(local)
        if rank == 0:
            ok = [i for i, _ in merged] == list(range(n_items))
            ok &= all(len(b) == i % 3 and all((bb == i * 10 + k).all() for k, bb in enumerate(b)) for i, (b, _) in merged)
            q.put(bool(ok))
        else:
            assert merged is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_results_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_results_single_process():
    local = [(3, ([], [])), (1, ([np.zeros((4, 2), np.int64)], [0.9]))]
    assert [i for i, _ in gather_results(local)] == [1, 3]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver launches like ours): under torchrun only rank 0 works and prints
    one JSON line with impl / cpu_baseline / e2e; every other rank exits 0 silently."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "images/sec" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["n_gpus"] == 2
    assert line["steps"] == 1 and line["warmup"] == 0  # the arm times exactly --steps K after --warmup W
