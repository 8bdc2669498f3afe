"""Repeats tests/test_trainer_step.py::test_trainer_loop_reduces_loss N times in one process and prints the trajectories
(is the 30-step Adam loop reproducible run to run?).  usage: python tools/probes/loop_repeat.py [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
mm = g.load_package()
import parity_utils as pu
import stand_in_encoder as se
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
DEV = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for rep in range(N):
    for tex_mirror in (False, True):
        B, size = 8, 64
        dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), size, image_weight=1.0)
        images = se.make_images(B, size, size, 43).to(DEV)
        enc = se.make_encoder(dr.vertices_init, size, size, 42, tex_mirror=tex_mirror).to(DEV)
        opt = torch.optim.Adam(enc.parameters(), lr=2e-3, betas=(0.5, 0.999))
        losses = []
        for _ in range(30):
            opt.zero_grad(set_to_none=True)
            loss, _ = se.trainer_step_loss(dr, enc, images)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        print(rep, tex_mirror, " ".join("%.4f" % l for l in losses[:8]), "...", " ".join("%.4f" % l for l in losses[-5:]), flush=True)
