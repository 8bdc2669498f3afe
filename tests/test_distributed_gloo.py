"""world_size-2 gloo test (CPU) of the multi-GPU plumbing in bench.py: the render path shards over images with NO
data-path collective; the only cross-rank step is the timing / throughput aggregation."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"], os.environ["LOCAL_RANK"], os.environ["WORLD_SIZE"] = str(rank), str(rank), str(world)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    assert bench.dist_env() == (rank, rank, world)
    # rank 1 is slower; both processed 48 images/step for 10 steps
    ms, units = bench.aggregate(12.0 + 3.0 * rank, 48 * 10, world)
    seeds = [bench.shard_seed(r) for r in range(world)]
    if rank == 0:
        torch.save({"ms": ms, "units": units, "seeds": seeds}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_takes_max_time_and_sum_units(tmp_path):
    out = str(tmp_path / "r.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ms"] == 15.0                # MAX over ranks (device-timed), never a mean
    assert r["units"] == 960.0            # whole-job units = SUM over ranks -> weak scaling
    assert len(set(r["seeds"])) == 2      # every rank draws its own shard of the synthetic data


def test_single_process_aggregate_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.aggregate(3.5, 100, 1) == (3.5, 100)


def test_algorithmic_bytes_match_survey_table():
    """cfg-2 = 1.137 MB/img fwd, 1.719 MB/img bwd (with the upstream-grad term), 137.2 MB batch."""
    sys.path.insert(0, ROOT)
    import bench
    fwd, bwd, step = bench.algorithmic_bytes(48, 642, 1280, 128, 128, 256, 128, bg=True, extra=True)
    assert abs(fwd / 48 / 1e6 - 1.137) < 2e-3
    assert abs(bwd / 48 / 1e6 - 1.719) < 2e-3
    assert abs(step / 1e6 - 137.2) < 0.2
