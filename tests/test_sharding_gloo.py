"""N>1 path of bench.py on CPU: world_size-2 gloo run of the sharding / timing-reduction logic
(frames are sharded across ranks with no data-path collective; the only exchange is max-over-ranks)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    seeds = bench.shard_seeds(rank, 4)
    ms = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)
    red = bench.reduce_max(ms, dist)
    allseeds = [None] * world
    dist.all_gather_object(allseeds, seeds)
    if rank == 0:
        print(json.dumps({"seeds": allseeds, "max": red.tolist(), "value": bench.whole_job_mps(world, 4, 10, red[0].item())}))
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_sharding_and_max_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["seeds"] == [[0, 1, 2, 3], [4, 5, 6, 7]]                  # disjoint frames per rank
    assert d["max"] == [11.0, 20.0]                                      # max over ranks
    assert abs(d["value"] - 2 * 4 * 10 * 8.2944 / 0.011) < 1e-6          # whole-job MP/s from the slowest rank
