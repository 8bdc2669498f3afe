#!/usr/bin/env python
"""bench.py -- throughput of the fused render + recon_data forward+backward (the hot path).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun by the driver)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port; rank 0 only)

A "step" is one pass of mm_render_compare_fwd_bwd over one synthetic batch of BASELINE.json configs[1]
(B=48 per GPU, ellipsoid template V=642 F=1280, 128x128, texture 256x128, no_mask, contour 0.1).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "train images/sec (fused render+loss fwd+bwd) @128x128"
UNIT = "images/s"
B_PER_GPU = 48
IMAGE_SIZE = 128
NSETS = 4                     # rotating input/output sets: 4 x ~137 MB > 126 MB L2, so every step starts L2-cold
# intervals between the CUDA events the library records around its launch groups (mm_ctx_set_timing).  Recording the
# events switches programmatic dependent launch off across them, so these figures are a little above what the same
# kernels cost inside the timed step; they are used for the roofline of the dominant kernel only.
KERNELS = ["vertex_fwd", "geom_fwd", "shade_fused", "gsoft", "geom_bwd", "vertex_bwd", "tail"]
# kernels of one fused step (mm_ctx_get_int "fused_kernels"): k_vertex_fwd, k_scatter_hard, k_soft_fwd, k_shade<fused> (which
# also re-does the truncated pixels), k_soft_bwd (truncated pixels + pair list), k_vertex_bwd (which also finalises the loss);
# k_gsoft only runs for H or W not a multiple of 4

def algorithmic_bytes(B, V, F, H, W, Ht, Wt, bg=True, extra=False):
    """bytes each tensor contributes when touched once per direction (fp32)."""
    fwd = 4 * (3 * V + 3 * Ht * Wt + (3 * H * W if bg else 0) + 4 * H * W + 14) + 4 * (4 * H * W + 3 * F)
    bwd = 4 * ((4 * H * W if extra else 0) + 3 * Ht * Wt + (3 * H * W if bg else 0) + 4 * H * W + 3 * V + 14) + \
        4 * (3 * V + 3 * Ht * Wt + (3 * H * W if bg else 0) + 14)
    return B * fwd + F * 36, B * bwd, B * (fwd + bwd) + F * 36


# ------------------------------------------------------------------------------------------ distributed helpers
def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def aggregate(local_ms, local_units, world):
    """Whole-job figures: time = MAX over ranks (device-timed), units = SUM over ranks."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return local_ms, local_units
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def shard_seed(rank):
    """Every rank draws its own shard of the synthetic data set (manual_seed(1234 + rank))."""
    return 1234 + rank


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        if self.nv is None:
            return

        def loop():
            while not self._stop.is_set():
                self._once()
                time.sleep(0.002)
        self._once()
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        self._once()
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ workload
def build_workload(mm, device, rank, nsets=NSETS, B=B_PER_GPU):
    import torch
    import parity_utils as pu
    dr = mm.DiffRender(pu.get_mesh(mm, "ellipsoid"), IMAGE_SIZE, ratio=1, init_ellipsoid=1, image_weight=1.0)
    H = W = IMAGE_SIZE
    sets = []
    for i in range(nsets):
        A = pu.make_attributes(dr.vertices_init, B, H, W, shard_seed(rank) * 100 + i)
        G = pu.make_attributes(dr.vertices_init, B, H, W, shard_seed(rank) * 100 + 50 + i)
        sets.append((A, G))
    return dr, sets


class FusedRunner(object):
    """Pre-allocates everything and calls mm_render_compare_fwd_bwd directly (inputs resident in HBM)."""

    def __init__(self, mm, dr, sets_cpu, device, tex_mirror=False):
        import torch
        self.torch, self.mm, self.dr, self.dev = torch, mm, dr, device
        self.tex_mirror = tex_mirror      # SURVEY 8(f)-3 variant: the texture's upper half only (not the headline workload)
        self.L = mm.lib()
        self.h = dr._ctx(torch.device(device))
        self.sets = []
        B = sets_cpu[0][0]['azimuths'].shape[0]
        self.B = B
        H, W, F = dr.height, dr.image_size, dr.num_faces
        for A, G in sets_cpu:
            d = {k: v.to(device).contiguous() for k, v in A.items()}
            if tex_mirror:
                d['textures'] = d['textures'][:, :, :d['textures'].shape[2] // 2].contiguous()
            with torch.no_grad():
                gt, _ = dr.render(no_mask=True, **{k: v.to(device) for k, v in G.items()})
            d['gt'] = gt.contiguous()
            d['out'] = {
                'rgba': torch.empty(B, 4, H, W, device=device), 'fn': torch.empty(B, F, 3, device=device),
                'loss': torch.empty(4, device=device), 'g_v': torch.empty_like(d['vertices']),
                'g_az': torch.empty(B, device=device), 'g_el': torch.empty(B, device=device),
                'g_di': torch.empty(B, device=device), 'g_bi': torch.empty(B, 2, device=device),
                'g_tex': torch.empty_like(d['textures']), 'g_li': torch.empty(B, 9, device=device),
                'g_bg': torch.empty_like(d['bg']), 'ws': self.h.workspace(B)}
            self.sets.append(d)
        self.stream = torch.cuda.current_stream()

    def step(self, i, contour=0.1):
        d = self.sets[i % len(self.sets)]
        o = d['out']
        p = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
        Ht, Wt = d['textures'].shape[2] * (2 if self.tex_mirror else 1), d['textures'].shape[3]
        rc = self.L.mm_render_compare_fwd_bwd(
            self.h.handle, self.B, p(d['vertices']), p(d['azimuths']), p(d['elevations']), p(d['distances']),
            p(d['biases']), p(d['textures']), Ht, Wt, 1 if self.tex_mirror else 0, p(d['lights']), p(d['bg']), 1, p(d['gt']),
            1.0, contour, 1.0, ctypes.c_void_p(0), ctypes.c_void_p(0), p(o['rgba']), p(o['fn']), p(o['loss']),
            p(o['g_v']), p(o['g_az']), p(o['g_el']), p(o['g_di']), p(o['g_bi']), p(o['g_tex']), p(o['g_li']),
            p(o['g_bg']), p(o['ws']), o['ws'].numel(), ctypes.c_void_p(self.stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(self.L.mm_last_error().decode())
        return o


class E2ERunner(object):
    """The reference-facing path a trainer.py user takes -- DiffRender.render -> recon_data -> backward()
    (trainer.py:276,441,509) -- fed from pinned HOST buffers every step, loss read back to the host.
    The host->device copies of step i+1 are issued on a copy stream while step i computes (double-buffered device
    staging, the same overlap the reference gets from its DataLoader + `.cuda(non_blocking=True)`, trainer.py:247);
    every byte is still copied inside the timed region, every step."""

    def __init__(self, mm, dr, sets_cpu, device, tex_mirror=False):
        import torch
        self.torch, self.dr, self.dev = torch, dr, device
        self.tex_mirror = tex_mirror          # labelled variant: the un-concatenated half of the atlas crosses PCIe
        keys = ['vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures', 'lights', 'bg']
        self.keys = keys
        # one pinned block per batch (what a DataLoader's collate + pin_memory hands over) and one device staging block per
        # slot: a step's inputs cross PCIe as ONE copy instead of nine; the tensors the API sees are views into the block
        shapes = None
        self.host = []
        for A, G in sets_cpu:
            with torch.no_grad():
                gt, _ = dr.render(no_mask=True, **{k: v.to(device) for k, v in G.items()})
            src = {k: A[k].contiguous().float() for k in keys}
            if tex_mirror:
                src['textures'] = src['textures'][:, :, :src['textures'].shape[2] // 2].contiguous()
            src['gt'] = gt.cpu().contiguous()
            if shapes is None:
                shapes, off = {}, 0
                for k, v in src.items():
                    shapes[k] = (off, v.numel(), tuple(v.shape))
                    off += (v.numel() + 63) // 64 * 64           # 256-byte aligned views
                self.total = off
            flat = torch.empty(self.total, dtype=torch.float32).pin_memory()
            for k, v in src.items():
                o, n, _ = shapes[k]
                flat[o:o + n].copy_(v.reshape(-1))
            self.host.append(flat)
        self.shapes = shapes
        self.h2d_bytes = sum(n for _, n, _ in shapes.values()) * 4
        self.loss_host = torch.empty(1).pin_memory()
        self.d2h_bytes = 4
        self.copy_stream = torch.cuda.Stream(device=device)
        self.stage_flat = [torch.empty(self.total, dtype=torch.float32, device=device) for _ in range(2)]
        self.stage = [{k: f[o:o + n].view(shp) for k, (o, n, shp) in shapes.items()} for f in self.stage_flat]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]       # staging buffer filled
        self.free = [torch.cuda.Event(), torch.cuda.Event()]        # staging buffer consumed
        self.primed = False

    def _upload(self, i):
        torch = self.torch
        slot = i % 2
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            self.stage_flat[slot].copy_(self.host[i % len(self.host)], non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def step(self, i):
        torch = self.torch
        cur = torch.cuda.current_stream()
        if not self.primed:
            for e in self.free:
                e.record(cur)
            self._upload(i)
            self.primed = True
        self._upload(i + 1)                                   # next step's inputs, overlapped with this step's compute
        slot = i % 2
        cur.wait_event(self.ready[slot])
        st = self.stage[slot]
        A = {k: st[k].detach().requires_grad_(True) for k in self.keys}
        if self.tex_mirror:
            A['_tex_mirror'] = True
        rgbs, _ = self.dr.render(no_mask=True, **A)
        loss = self.dr.recon_data(rgbs, st['gt'], no_mask=True, contour=0.1)
        loss.backward()
        self.free[slot].record(cur)
        self.loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss


class ApiRunner(object):
    """The reference's three calls -- render (trainer.py:276), recon_data (:441), backward (:509) -- with device-resident
    inputs: what a trainer.py user gets from the drop-in without any patch (lazy fusion: recon_data's gradient is formed
    inside the render backward).  `graph=True` captures one step per input set into a CUDA graph and replays it (the
    library never allocates or synchronises, so the whole step is capturable; this is how a launch-bound inner loop is meant
    to be driven on this hardware) -- the eager figure is bounded by Python / autograd dispatch (~0.2 ms of host time per step)."""

    def __init__(self, mm, dr, fused, graph):
        import torch
        self.torch, self.dr, self.graph = torch, dr, graph
        self.keys = ['vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures', 'lights', 'bg']
        self.sets = fused.sets
        self.graphs = []
        self.last = None
        if graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                  # warm-up off the capturing stream (PyTorch graph-capture protocol)
                for d in self.sets:
                    self._eager(d)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for d in self.sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep = self._eager(d)
                self.graphs.append((g, keep))

    def _eager(self, d):
        A = {k: d[k].detach().requires_grad_(True) for k in self.keys}
        rgbs, _ = self.dr.render(no_mask=True, **A)
        loss = self.dr.recon_data(rgbs, d['gt'], no_mask=True, contour=0.1)
        loss.backward()
        return loss, A, rgbs

    def step(self, i):
        if self.graph:
            g, keep = self.graphs[i % len(self.graphs)]
            g.replay()
            self.last = keep
        else:
            self.last = self._eager(self.sets[i % len(self.sets)])
        return self.last


def pcie_ceiling(torch, device, nbytes, reps=30, trials=3):
    """Pinned host -> device copy of one e2e-sized block, alone on the link (best of `trials`): the ceiling the e2e line is
    compared with."""
    src = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    src.fill_(1.0)                                   # touch every page
    dst = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
    best = 0.0
    for _ in range(trials):
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def timed(torch, world, fn, steps):
    """barrier + sync | start event | `steps` calls | end event | sync + barrier; returns local ms."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return e0.elapsed_time(e1)


def cfg1_cpu_rows(mm, reps=10):
    """BASELINE.json configs[0] / BASELINE.md section 5 rows C1, C2, C4: the CPU restatement on ONE image of sphere.obj at
    64x64 (azim 30, elev 15, dist 4.5, texture 128x64, GT mask = centred disc r = 0.45 W), soft-IoU against the disc.
    C1 = forward + soft IoU, C2 = forward + backward; median of `reps`, all host threads of the OpenMP kernels."""
    import torch
    import parity_utils as pu
    dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 64, ratio=1, init_ellipsoid=1, image_weight=1.0)
    orc = pu.oracle_for(dr)
    g = torch.Generator().manual_seed(1234)
    A = {'azimuths': torch.tensor([30.0]), 'elevations': torch.tensor([15.0]), 'distances': torch.tensor([4.5]),
         'biases': torch.zeros(1, 2), 'vertices': dr.vertices_init[None].clone(), 'textures': torch.rand(1, 3, 128, 64, generator=g),
         'lights': torch.tensor([[3.0] + [0.0] * 8]), 'bg': torch.rand(1, 3, 64, 64, generator=g)}
    yy, xx = torch.meshgrid(torch.arange(64.0), torch.arange(64.0), indexing='ij')
    disc = (((xx - 31.5) ** 2 + (yy - 31.5) ** 2) < (0.45 * 64) ** 2).float()
    gt = torch.cat([torch.rand(1, 3, 64, 64, generator=g), disc[None, None]], 1)

    def fwd():
        with torch.no_grad():
            rgb = orc.render(no_mask=False, **A)[0]
            return float(orc.recon_data(rgb, gt, contour=0)), float((rgb[:, 3] * disc).sum() / ((rgb[:, 3] + disc - rgb[:, 3] * disc).sum() + 1e-10))

    def fwdbwd():
        Ag = {k: v.clone().requires_grad_(True) for k, v in A.items()}
        rgb = orc.render(no_mask=False, **Ag)[0]
        orc.recon_data(rgb, gt, contour=0).backward()

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]
    loss, iou = fwd()
    t1, t2 = med(fwd), med(fwdbwd)
    return {"config": "cfg-1: sphere.obj V=642 F=1280, 1 image, 64x64, tex 128x64, azim 30 elev 15 dist 4.5, GT mask = disc r=0.45W",
            "soft_iou_vs_disc": iou, "loss": loss, "C1_fwd_ms": 1e3 * t1, "C1_images_per_s": 1.0 / t1,
            "C2_fwd_bwd_ms": 1e3 * t2, "C2_images_per_s": 1.0 / t2, "reps": reps}


def cpu_baseline(mm, dr, sets_cpu, budget_s=15.0):
    """Oracle port (our CPU restatement of the Kaolin DIB-R path + torch-CPU glue) on a bounded sample: whole batches of the
    bench workload (B=48), as many as fit the time budget; plus the cfg-1 rows."""
    import torch
    import parity_utils as pu
    orc = pu.oracle_for(dr)
    A, G = sets_cpu[0]
    nb = A['azimuths'].shape[0]
    with torch.no_grad():
        gt, _, _, _ = orc.render(no_mask=True, **G)

    def once():
        Ag = {k: v.clone().requires_grad_(k != 'delta_vertices') for k, v in A.items()}
        rgb, _, _, _ = orc.render(no_mask=True, **Ag)
        orc.recon_data(rgb, gt, no_mask=True, contour=0.1).backward()
    once()
    t0 = time.perf_counter()
    once()
    one = time.perf_counter() - t0
    reps = max(1, min(200, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = time.perf_counter() - t0
    return {"value": nb * reps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d whole batches (B=%d) of the bench workload, fwd+bwd, %.1f s; oracle = our CPU restatement of "
                      "Kaolin DIB-R (kaolin itself is CUDA-only and not installable offline), host has %d logical cores"
                      % (reps, nb, dt, os.cpu_count()),
            "cfg1": cfg1_cpu_rows(mm)}


def run_reference(args):
    """--impl reference: the CPU implementation of the same path (oracle port), rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    import __graft_entry__ as g
    g.build_oracle()
    mm = g.load_package()
    dr, sets = build_workload(mm, "cpu", 0, nsets=1, B=B_PER_GPU)
    import parity_utils as pu
    orc = pu.oracle_for(dr)
    A, G = sets[0]
    nb = B_PER_GPU
    with torch.no_grad():
        gt, _, _, _ = orc.render(no_mask=True, **G)

    def once():
        Ag = {k: v.clone().requires_grad_(k != 'delta_vertices') for k, v in A.items()}
        rgb, _, _, _ = orc.render(no_mask=True, **Ag)
        orc.recon_data(rgb, gt, no_mask=True, contour=0.1).backward()
    steps = min(args.steps, 20)                     # ~1 s per B=48 batch on 16 host threads: bounded to a few minutes
    for _ in range(min(args.warmup, 2)):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    v = nb * steps / dt
    sample = "each step = one whole batch (B=%d) of the bench workload, fwd+bwd, %d steps (of --steps %d)" % (nb, steps, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg-2 (BASELINE.json configs[1]): B=%d, ellipsoid V=642 F=1280, 128x128, tex 256x128, "
                                   "no_mask, contour 0.1, render+recon_data fwd+bwd; each step = one whole batch on the CPU" % nb},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4-ddp"],
                    help="cfg2 (default): the headline render-compare step; cfg3 / cfg4-ddp: a full trainer step (tools/trainer_bench.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="timed region only (for ncu): no e2e / cpu_baseline / per-kernel pass")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "cfg2":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import trainer_bench
        return trainer_bench.run(args, sys.modules[__name__])

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    if world > 1:
        # one process per GPU: keep this rank's threads (and the first touch of its pinned host buffers) on the CPUs next to
        # its GPU, so that N ranks do not pull their e2e uploads across the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the ONE JSON line only: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION (the log file
        # setting is honoured above that level only), and its INFO / WARN output to stdout unless a file is named
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(device))
    if not os.path.exists(os.path.join(g.PKG_DIR, "libmagicmirror.so")):
        g.build_cuda()
    mm = g.load_package()
    K, Wm = args.steps, max(args.warmup, 3)

    dr, sets = build_workload(mm, device, rank)
    fused = FusedRunner(mm, dr, sets, device)
    V, F, H, W = dr.num_vertices, dr.num_faces, dr.height, dr.image_size
    Ht, Wt = sets[0][0]['textures'].shape[2:]
    bytes_fwd, bytes_bwd, bytes_step = algorithmic_bytes(B_PER_GPU, V, F, H, W, Ht, Wt)

    # ---- value: device-resident inputs, exactly K timed steps
    for i in range(Wm):
        fused.step(i)
    clk = ClockSampler(local_rank)
    clk.start()
    ms_local = timed(torch, world, fused.step, K)
    clocks = clk.stop()
    ms, units = aggregate(ms_local, B_PER_GPU * K, world)
    value = units / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events recorded by the library around each of its launches)
    kms = None
    L = mm.lib()
    if hasattr(L, "mm_ctx_set_timing") and not args.profile:
        h = fused.h.handle
        L.mm_ctx_set_timing(h, 1)
        acc = [0.0] * len(KERNELS)
        buf = (ctypes.c_float * 8)()
        nprof = min(K, 200)
        for i in range(nprof):
            fused.step(i)
            L.mm_ctx_get_timing(h, buf, 8)
            for j in range(len(KERNELS)):
                acc[j] += buf[j]
        L.mm_ctx_set_timing(h, 0)
        kms = {k: acc[j] / nprof for j, k in enumerate(KERNELS)}

    if args.profile:
        if os.environ.get("MM_PROFILE_API"):            # ncu: also a few eager steps of the three-call API path
            r = ApiRunner(mm, dr, fused, graph=False)
            for i in range(4):
                r.step(i)
            torch.cuda.synchronize()
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "ms_per_step": ms / K}), flush=True)
        return
    fused_kernels = L.mm_ctx_get_int(fused.h.handle, b"fused_kernels")

    # ---- value_api: the reference's three calls (render -> recon_data -> backward) with device-resident inputs, eager and as
    # replayed CUDA graphs (same kernels either way; the eager figure is bounded by Python / autograd host time)
    api = {}
    for mode in ("graph", "eager"):
        try:
            r = ApiRunner(mm, dr, fused, graph=(mode == "graph"))
            Ka = max(3, min(K, 500))
            for i in range(Wm):
                r.step(i)
            ms_a_local = timed(torch, world, r.step, Ka)
            ms_a, units_a = aggregate(ms_a_local, B_PER_GPU * Ka, world)
            api[mode] = {"value": units_a / (ms_a * 1e-3), "ms_per_step": ms_a / Ka, "steps": Ka}
            del r
        except Exception as e:                      # a capture failure must not take the headline numbers down with it
            api[mode] = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- e2e: reference-facing API, pinned host inputs copied every step, loss read back
    e2e_runner = E2ERunner(mm, dr, sets, device)
    Ke = max(3, min(K, 200))
    for i in range(3):
        e2e_runner.step(i)
    torch.cuda.synchronize()
    e2e_runner.primed = False                 # the timed region uploads its own first batch
    ms_e_local = timed(torch, world, e2e_runner.step, Ke)
    ms_e, units_e = aggregate(ms_e_local, B_PER_GPU * Ke, world)
    h2d_rank = e2e_runner.h2d_bytes * Ke / (ms_e_local * 1e-3) / 1e9          # this rank's achieved host->device rate
    pcie_alone = pcie_ceiling(torch, device, e2e_runner.h2d_bytes) if rank == 0 else 0.0
    # labelled variant, NOT the headline: the texture handed over as the un-concatenated half (render(_tex_mirror=True))
    ev = E2ERunner(mm, dr, sets, device, tex_mirror=True)
    for i in range(3):
        ev.step(i)
    torch.cuda.synchronize()
    ev.primed = False
    ms_v, units_v = aggregate(timed(torch, world, ev.step, Ke), B_PER_GPU * Ke, world)
    e2e_mirror = {"value": units_v / (ms_v * 1e-3), "h2d_bytes_per_step": ev.h2d_bytes,
                  "what": "same three calls, atlas crosses PCIe as the un-concatenated half (bit-identical image); not the headline"}
    del ev
    h2d_all = None
    if world > 1:
        t = torch.tensor([h2d_rank], dtype=torch.float64, device=device)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        h2d_all = [float(x.item()) for x in gathered]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src, "traffic": None}
        Bq = B_PER_GPU
        if kms:
            # dominant kernel: k_shade<FUSED> (shading forward + loss sums + the whole RGB-side backward).  THREE byte counts:
            #  contract     SURVEY 8(d): every image-sized tensor of the step once PER DIRECTION (tex, bg, gt count twice although
            #               this kernel reads them once) -- the figure `achieved` / `frac` must be computed from;
            #  single_touch the API tensors this launch actually reads (tex, bg, gt) and writes (rgba, g_bg, g_tex), once each;
            #  dram         what crossed the DRAM pins during the launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the
            #               committed capture, profiles/traffic.json written by tools/ncu_traffic.py) -- below the algorithmic
            #               figure because written planes stay in the 126 MB L2.
            nbytes = bytes_step - Bq * 4 * (3 * V + 3 * F + 3 * V + 3 * V + 14 * 3) - F * 36
            single = Bq * 4 * (2 * 3 * Ht * Wt + 2 * 3 * H * W + 2 * 4 * H * W)
            t_dom = kms["shade_fused"]
            ach = nbytes / (t_dom * 1e-3) / 1e9
            roof.update({"kernel": "k_shade<fused>", "achieved": ach, "frac": ach / peak,
                         "algorithmic_bytes_per_launch": nbytes, "avg_launch_ms": t_dom, "kernel_ms": kms,
                         "frac_contract": ach / peak,
                         "frac_single_touch": single / (t_dom * 1e-3) / 1e9 / peak, "single_touch_bytes_per_launch": single})
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                roof["traffic"] = tr.get("k_shade_fused", {}).get("dram_bytes_per_launch")
                roof["traffic_source"] = tr.get("k_shade_fused", {}).get("source")
                if roof["traffic"]:
                    roof["frac_dram"] = roof["traffic"] / (t_dom * 1e-3) / 1e9 / peak
            except Exception:
                pass
        ach_step = bytes_step / (ms / K * 1e-3) / 1e9
        roof["step"] = {"achieved": ach_step, "frac": ach_step / peak, "algorithmic_bytes_per_step": bytes_step,
                        "note": "whole fused step against the HBM peak: THE figure north_star's 70 % target is about"}
        e2e_val = units_e / (ms_e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg-2 (BASELINE.json configs[1]): B=%d/GPU, ellipsoid V=%d F=%d, %dx%d, tex %dx%d, "
                                   "no_mask, contour 0.1, fused render+recon_data fwd+bwd" % (B_PER_GPU, V, F, H, W, Ht, Wt),
                       "l2": "inputs larger than L2: %d rotating input/output sets (%d x %.0f MB)" % (NSETS, NSETS, bytes_step / 1e6),
                       "parallelism": "dp%d (images sharded, no data-path collective)" % world},
            "clocks": clocks,
            "value_api": api.get("graph", {}).get("value"),
            "api": {"what": "DiffRender.render -> recon_data -> backward() (trainer.py:276,441,509), inputs resident in HBM, "
                            "lazy fusion; 'graph' = the same three calls captured once per input set and replayed",
                    "graph": api.get("graph"), "eager": api.get("eager"),
                    "kernels_per_step": L.mm_ctx_get_int(fused.h.handle, b"api_kernels")},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": e2e_runner.h2d_bytes,
                    "d2h_bytes_per_step": e2e_runner.d2h_bytes, "steps": Ke,
                    "h2d_gbs_per_rank": h2d_all if h2d_all else [h2d_rank],
                    "h2d_gbs_aggregate": sum(h2d_all) if h2d_all else h2d_rank,
                    "pcie_h2d_ceiling_gbs": pcie_alone,
                    "variant_mirrored_atlas": e2e_mirror,
                    "pcie_note": "ceiling = the same pinned block copied alone on rank 0's link; e2e is bound by it "
                                 "(compute per step is %.3f ms, the copy %.3f ms)" % (ms / K, e2e_runner.h2d_bytes / max(pcie_alone, 1e-9) / 1e6),
                    "api": "DiffRender.render -> recon_data -> backward; one pinned host block per batch copied every step on a copy stream (double-buffered)"},
            "gpu_launches": fused_kernels * K,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            g.build_oracle()
            line["cpu_baseline"] = cpu_baseline(mm, dr, sets)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
