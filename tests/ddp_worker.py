"""TEST INFRASTRUCTURE: one rank of the data-parallel trainer step: images sharded over ranks, the render path runs
per rank with no collective, DistributedDataParallel all-reduces (averages) the encoder gradients -- NCCL when every rank
has its own GPU, gloo when the ranks share cuda:0 (the single-GPU test box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_rank(rank, world, port, B_per_rank, size, seed, out_path, mesh_name, sized=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import __graft_entry__ as g
    import parity_utils as pu
    import stand_in_encoder as se
    mm = g.load_package()
    own_gpu = torch.cuda.device_count() >= world
    dev = torch.device("cuda", rank if own_gpu else 0)
    torch.cuda.set_device(dev)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl" if own_gpu else "gloo", rank=rank, world_size=world)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dr = mm.DiffRender(pu.get_mesh(mm, mesh_name), size, image_weight=1.0)
    if sized:          # the reference's parameter volume (33.8 M, tools/sized_encoder.py), BatchNorm frozen so that shards == whole batch
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sized_encoder
        torch.manual_seed(seed)
        enc = sized_encoder.SizedEncoder(dr, dr.height, dr.image_size, amp=False).to(dev).eval()
    else:
        enc = se.make_encoder(dr.vertices_init, dr.height, dr.image_size, seed).to(dev)
    ddp = DDP(enc, device_ids=[dev.index])
    images = se.make_images(B_per_rank * world, dr.height, dr.image_size, seed + 1)
    shard = images[rank * B_per_rank:(rank + 1) * B_per_rank].to(dev)       # rank r takes images [r*B, (r+1)*B)
    loss, _ = se.trainer_step_loss(dr, ddp, shard)
    loss.backward()                                                          # DDP: all-reduce(sum) / world of every gradient
    torch.cuda.synchronize()
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    if rank == 0:
        torch.save({"grads": {n: p.grad.detach().cpu() for n, p in enc.named_parameters()},
                    "loss_mean": float(lsum) / world, "backend": dist.get_backend()}, out_path)
    dist.barrier()
    dist.destroy_process_group()
