"""World-size-2 CPU (gloo) tests of the multi-process logic of bench.py: image sharding, max-over-ranks
timing reduction, and the reference arm under torchrun (rank 0 prints one JSON line, other ranks are silent)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.shard_images(8, rank, world)
    got = [None] * world
    dist.all_gather_object(got, mine)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # the bench reports the slowest rank's time
    dist.barrier()
    if rank == 0:
        ret["shards"] = got
        ret["max"] = float(t)
        ret["env"] = bench.dist_env()
    dist.destroy_process_group()


def test_shard_and_max_over_ranks_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, ret), nprocs=2, join=True)
    assert ret["shards"] == [[0, 1, 2, 3], [4, 5, 6, 7]]          # engine.py:86-88: batch_size // world_size images per rank
    assert ret["max"] == 11.0
    assert ret["env"] == (0, 0, 2)


def test_alg_bytes_match_survey():
    import bench
    assert bench.alg_bytes(8, 512, 97, 97, 4, True, False) == 346853376     # 346.85 MB / step
    assert bench.alg_bytes(8, 512, 97, 97, 4, False, True) == 693706752
    assert bench.alg_bytes(8, 512, 97, 97, 2, True, True) == (346853376 + 693706752) // 2


def test_reference_arm_under_torchrun_world2():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "cca_fwd_bwd_pixels_per_s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0
