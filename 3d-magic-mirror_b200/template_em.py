"""Per-epoch template update of the reference (`trainer.py:994-1105`, the "EM" step) for a data-parallel run.

SURVEY 8(e)-3: the only cross-rank step on this path besides the gradient all-reduce.  Every rank renders its shard of the
training set (`_, Ae0 = diffRender.render(**Ae)`, trainer.py:1005) and accumulates `delta_vertices`; the reference then
averages over the WHOLE set, smooths with the uniform Laplacian, clips, and moves `vertices_init` -- rolling back if a vertex
crossed the depth-sign plane.  Sharded: one all-reduce of a (V,3) sum + a count (7.7 KB at V=642), after which every rank
applies the same deterministic update -- equivalent to computing on rank 0 and broadcasting `vertices_init`, without the
second collective.  Host-side torch (V-sized, once per epoch): no kernel.
"""
import torch


def template_update(vertices_init, sum_delta, count, laplacian, em_step=1.0, warm_up=1.0, smooth=0.0, extra_smooth=0,
                    clip=0.05, white=False, cross=True):
    """trainer.py:1071-1100 for the 'all average' rule (opt.em == 1; the sample-selection variants only change which samples
    enter `sum_delta` / `count`).  vertices_init (1,V,3) or (V,3); sum_delta (V,3) = sum over samples of delta_vertices.
    Returns (new_vertices_init, updated: bool, whether_cross)."""
    v0 = vertices_init.reshape(-1, 3)
    if count <= 1:
        return vertices_init, False, 0.0
    last = sum_delta.to(v0) * 1.0 / count                                          # :1073
    if smooth > 0:                                                                   # :1074-1081
        lap = laplacian.to(v0)
        for _ in range(1 + int(extra_smooth)):
            last = last + torch.matmul(lap, last) * smooth
    last = last.clamp(-clip, clip)                                                   # :1082-1083
    new = v0 + warm_up * em_step * last                                              # :1084
    if white:
        new = new - new.mean(dim=0, keepdim=True)                                    # :1088-1089
    whether_cross = float(torch.relu(-torch.sign(new[:, 2]) * torch.sign(v0[:, 2])).sum())      # :1092
    if whether_cross > 0 and cross:                                                  # :1095-1096: keep the old template
        return vertices_init, False, whether_cross
    return new.reshape(vertices_init.shape), True, whether_cross


def sharded_template_update(vertices_init, local_sum_delta, local_count, laplacian, group=None, **kw):
    """Each rank passes the sum of `delta_vertices` over ITS shard of the training set and the number of samples in it.
    One all-reduce (sum) of V*3 + 1 floats; every rank then computes the identical update (fp32 sums of identical operands in
    identical order), so `vertices_init` stays bit-equal across ranks without a broadcast."""
    import torch.distributed as dist
    buf = torch.cat([local_sum_delta.reshape(-1).float(), torch.tensor([float(local_count)], device=local_sum_delta.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    total = buf[:-1].reshape(-1, 3)
    count = int(round(float(buf[-1])))
    return template_update(vertices_init, total, count, laplacian, **kw)
