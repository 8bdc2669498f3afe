"""bench.py --workload cfg3 | cfg4-ddp : one full trainer step around the render path (BASELINE.json configs[2] / configs[3]).

    cfg3      1 GPU, CUB shape: sphere.obj, 128x128, B=48 -- the reconstruction part of trainer.py:239-531: encode, render #1,
              the two novel-view renders (#2/#3, one pass through render_many), re-encode + render #4 (vertex stage only),
              recon_data + mesh regularisers + attribute-cycle loss, backward, Adam.
    cfg4-ddp  N GPUs, Market shape: smpl_uv_642.obj, ratio 2 -> 256x128, B=48 per rank -- encode, render, recon_data +
              regularisers, backward with DistributedDataParallel's bucketed NCCL all-reduce of the ~33.8 M fp32 gradients
              (135 MB) overlapped with it, Adam.  trainer.py:94-96 is the (broken) DataParallel this replaces.
The encoder is tools/sized_encoder.py (the reference's parameter volume and BatchNorm-per-rank; NOT its architecture).  The
discriminator and its losses are out of scope.  Besides the step time the line reports where it goes:
`encoder_ms` (encoder fwd+bwd+Adam alone), `render_path_ms` (render -> recon_data -> backward alone, eager),
`allreduce` (a stand-alone NCCL all-reduce of one flat fp32 buffer of the gradient volume: ms, algorithm and bus GB/s),
`step_nosync_ms` (the same step without the collective) -- all device-timed, max over ranks.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

B_PER_GPU = 48


def _timed(torch, dist, world, fn, steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run(args, bench):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    import parity_utils as pu
    import sized_encoder as se
    rank, local_rank, world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    mm = g.load_package()
    cfg4 = args.workload == "cfg4-ddp"
    if cfg4:
        dr = mm.DiffRender(pu.get_mesh(mm, "smpl_uv_642"), 128, ratio=2, init_ellipsoid=2, image_weight=1.0)
    else:
        dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 128, ratio=1, init_ellipsoid=1, image_weight=1.0)
    H, W, B = dr.height, dr.image_size, B_PER_GPU
    torch.manual_seed(1234)                                  # identical initial weights on every rank
    enc = se.SizedEncoder(dr, H, W).to(dev).to(memory_format=torch.channels_last)
    counts = enc.param_counts()
    model = torch.nn.parallel.DistributedDataParallel(enc, device_ids=[local_rank], gradient_as_bucket_view=True) if world > 1 else enc
    opt = torch.optim.Adam(enc.parameters(), lr=1e-4, betas=(0.5, 0.999), fused=True)
    import stand_in_encoder as sie
    nsets = 4
    images = [sie.make_images(B, H, W, bench.shard_seed(rank) * 10 + i).to(dev) for i in range(nsets)]

    def reg_loss(A):
        t = dr.regularizer_terms(A, temp=2.0)
        return dr.lambda_lpl * t['laplacian'] + dr.lambda_flat * t['flat'] + 0.1 * t['deform'] + 0.01 * t['depth'] + 0.1 * t['edge'] + 0.1 * t['flip']

    def step_cfg4(i, sync=True):
        X = images[i % nsets]
        ctx = model.no_sync() if (world > 1 and not sync) else _null()
        with ctx:
            Ae = model(X)
            Xer, Ae = dr.render(no_mask=True, **Ae)
            loss = dr.recon_data(Xer, X, no_mask=True, contour=0.1) + 0.1 * reg_loss(Ae)
            loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def step_cfg3(i, sync=True):
        X = images[i % nsets]
        Ae = model(X)
        Xer, Ae = dr.render(no_mask=True, **Ae)                                             # render #1 (trainer.py:276)
        # novel views (trainer.py:280-347): a rotated copy and a copy with attributes mixed across the batch
        Ae90 = dict(Ae); Ae90['azimuths'] = Ae['azimuths'] + 90.0
        perm = torch.roll(torch.arange(B, device=dev), 1)
        Ai = {k: (0.5 * (v + v[perm]) if k in ('vertices', 'delta_vertices', 'textures', 'lights') else v) for k, v in Ae.items()
              if torch.is_tensor(v)}
        Ai['bg'] = Ae['bg']
        (Xir, Ai), (Xer90, Ae90) = dr.render_many([Ai, Ae90], no_mask=True)                  # renders #2 / #3 in one pass
        Aire = model(Xir.detach())                                                           # trainer.py:365
        _, Aire = dr.render(no_mask=True, _need_image=False, **Aire)                         # render #4: vertex stage only (:367)
        l_cyc = sum(dr.recon_att(Aire, {k: v.detach() for k, v in Ai.items() if torch.is_tensor(v)}))
        loss = dr.recon_data(Xer, X, no_mask=True, contour=0.1) + 0.1 * (reg_loss(Ae) + reg_loss(Aire)) + 0.1 * l_cyc \
            + 1e-3 * (Xir.mean() + Xer90.mean())          # stands in for the adversarial term: a gradient into renders #2 / #3
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    step = step_cfg4 if cfg4 else step_cfg3
    K, Wm = min(args.steps, 200), max(min(args.warmup, 20), 3)
    for i in range(Wm):
        step(i)
    clk = bench.ClockSampler(local_rank)
    clk.start()
    ms = _timed(torch, dist, world, step, K)
    clocks = clk.stop()
    step_ms = ms / K

    # ---- where the step goes
    Kd = max(3, min(K, 30))
    detail = {}
    if world > 1:
        detail["step_nosync_ms"] = _timed(torch, dist, world, lambda i: step(i, sync=False), Kd) / Kd
        for p_ in enc.parameters():
            p_.grad = None
        n = sum(p_.numel() for p_ in enc.parameters())
        flat = torch.randn(n, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        t_ar = _timed(torch, dist, world, lambda i: dist.all_reduce(flat), 20) / 20
        nbytes = n * 4
        detail["allreduce"] = {"bytes": nbytes, "ms": t_ar, "algbw_gbs": nbytes / (t_ar * 1e-3) / 1e9,
                               "busbw_gbs": 2.0 * (world - 1) / world * nbytes / (t_ar * 1e-3) / 1e9,
                               "what": "stand-alone NCCL all-reduce of one flat fp32 buffer of the gradient volume"}

    def enc_only(i):
        A = model(images[i % nsets])
        l = sum(v.float().mean() for k, v in A.items() if torch.is_tensor(v))
        l.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    for i in range(3):
        enc_only(i)
    detail["encoder_ms"] = _timed(torch, dist, world, enc_only, Kd) / Kd
    with torch.no_grad():
        A0 = {k: v.detach() for k, v in enc(images[0]).items() if torch.is_tensor(v)}

    def render_only(i):
        # the render path of the step alone, on detached attributes: cfg-4 = render -> recon_data -> backward; cfg-3 = the four
        # renders of the step (render, render_many of two sets, vertex-stage-only render) + recon_data + regularisers -> backward
        A = {k: v.detach().requires_grad_(k != 'delta_vertices') for k, v in A0.items()}
        Xer, A = dr.render(no_mask=True, **A)
        loss = dr.recon_data(Xer, images[0], no_mask=True, contour=0.1)
        if not cfg4:
            A90 = dict(A); A90['azimuths'] = A['azimuths'] + 90.0
            (Xir, Ai), (X90, A90) = dr.render_many([dict(A), A90], no_mask=True)
            _, A4 = dr.render(no_mask=True, _need_image=False, **{k: v for k, v in A.items() if k not in ('face_normals', 'imnormal')})
            loss = loss + 0.1 * (reg_loss(A) + reg_loss(A4)) + 1e-3 * (Xir.mean() + X90.mean())
        loss.backward()
    for i in range(3):
        render_only(i)
    detail["render_path_ms"] = _timed(torch, dist, world, render_only, Kd) / Kd
    detail["render_path_share_of_step"] = detail["render_path_ms"] / step_ms
    detail["render_path_note"] = "eager Python dispatch included (the step is eager too); 4 renders per step" if not cfg4 else \
        "eager Python dispatch included (the step is eager too)"

    if rank == 0:
        line = {
            "metric": "train images/sec (full trainer step: encoder + render + loss + backward + Adam)", "value": world * B * K / (ms * 1e-3),
            "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 encoder (autocast, TF32 elsewhere) + f32 render path", "data": "synthetic",
            "config": {"workload": ("cfg-4 (BASELINE.json configs[3]): B=%d/GPU, smpl_uv_642 V=642 F=1280, %dx%d, tex %dx%d, DDP over %d ranks"
                                    if cfg4 else
                                    "cfg-3 (BASELINE.json configs[2]): B=%d, sphere V=642 F=1280, %dx%d, tex %dx%d, 4 renders per step, %d GPU")
                                   % (B, H, W, 2 * H, W, world),
                       "encoder": "tools/sized_encoder.py (parameter volume of the reference's AttributeEncoder, BatchNorm per rank; not its architecture)",
                       "encoder_params": counts, "parallelism": "ddp%d (NCCL bucketed all-reduce overlapped with backward)" % world},
            "clocks": clocks, "detail": detail,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
