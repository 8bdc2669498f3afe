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


def _em_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    import parity_utils as pu
    mm = g.load_package()
    dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 64)
    deltas = _em_deltas(dr.num_vertices)
    mine = deltas[rank::world]                                  # this rank's shard of the training set
    new, ok, cross = mm.sharded_template_update(dr.vertices_init[None], mine.sum(0), mine.shape[0],
                                                dr.vertices_laplacian_matrix, em_step=0.7, warm_up=0.5, smooth=0.3, clip=0.05)
    gathered = [torch.zeros_like(new) for _ in range(world)]
    dist.all_gather(gathered, new.contiguous())
    if rank == 0:
        torch.save({"new": new, "ok": ok, "cross": cross, "equal_across_ranks": all(torch.equal(gathered[0], t) for t in gathered)}, out)
    dist.barrier()
    dist.destroy_process_group()


def _em_deltas(V, n=37):
    g = torch.Generator().manual_seed(11)
    return 0.03 * torch.randn(n, V, 3, generator=g) + 0.01 * torch.randn(1, V, 3, generator=g)


def test_template_update_sharded_equals_single_process(tmp_path):
    """SURVEY 8(e)-3 / trainer.py:994-1105: the per-epoch template update over a sharded training set (uneven shards: 19 + 18
    samples) must give every rank the template a single process computes from the whole set."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    import parity_utils as pu
    mm = g.load_package()
    dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 64)
    deltas = _em_deltas(dr.num_vertices)
    want, ok, cross = mm.template_update(dr.vertices_init[None], deltas.sum(0), deltas.shape[0], dr.vertices_laplacian_matrix,
                                         em_step=0.7, warm_up=0.5, smooth=0.3, clip=0.05)
    assert ok and want.shape == (1, dr.num_vertices, 3) and not torch.equal(want[0], dr.vertices_init)
    # the reference's own lines (trainer.py:1073-1084), verbatim in form
    last = deltas.sum(0) * 1.0 / deltas.shape[0]
    last += torch.matmul(dr.vertices_laplacian_matrix, last) * 0.3
    last[last > 0.05] = 0.05
    last[last < -0.05] = -0.05
    assert torch.allclose(want[0], dr.vertices_init + 0.5 * 0.7 * last, atol=1e-7)
    out = str(tmp_path / "em.pt")
    port = 29500 + ((os.getpid() + 977) % 2000)
    mp.spawn(_em_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["equal_across_ranks"]
    assert torch.allclose(r["new"], want, atol=1e-6)            # the sum is re-associated across shards
    # a step that would push a vertex through the depth-sign plane is rolled back (trainer.py:1092-1096)
    big = torch.zeros(dr.num_vertices, 3)
    big[:, 2] = -torch.sign(dr.vertices_init[:, 2]) * 10.0
    same, ok2, cross2 = mm.template_update(dr.vertices_init[None], big * 4, 4, dr.vertices_laplacian_matrix, clip=5.0)
    assert not ok2 and cross2 > 0 and torch.equal(same[0], dr.vertices_init)
