#!/bin/bash
# usage (GPU box): tools/sanitize.sh  -- compute-sanitizer memcheck + racecheck over three small parity cases (fused, three-call lazy, materialised) + the glue kernels
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as g
mm = g.load_package()
import parity_utils as pu
for c in (dict(mesh="icosphere", B=2, image_size=32, no_mask=True, contour=0.1, seed=3),
          dict(mesh="sphere", B=2, image_size=22, ratio=1.5, no_mask=True, contour=0.1, seed=19),
          dict(mesh="sphere", B=1, image_size=48, no_mask=True, contour=0.0, seed=13, dist_range=(6.5, 7.0))):
    r = pu.run_parity_case(mm, **c)
    print("case ok", c["mesh"], r["face_idx_mismatch_staged"], r["loss_rel_err"], flush=True)
# round 2: the texture-flow kernels, the mesh regularisers and the template conditioning (fwd + bwd each)
import torch
dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 32, image_weight=1.0)
img = torch.rand(2, 3, 20, 12, device="cuda").requires_grad_(True)
flow = (torch.rand(2, 2, 17, 9, device="cuda") * 2.3 - 1.15).requires_grad_(True)
dr.texture_flow(img, flow, concat=True).sum().backward()
A = pu.to_device(pu.make_attributes(dr.vertices_init, 2, 32, 32, 5), "cuda:0", requires_grad=True)
_, out = dr.render(no_mask=True, _need_image=False, **A)
t = dr.regularizer_terms({'delta_vertices': A['delta_vertices'], 'vertices': A['vertices'], 'face_normals': out['face_normals']})
sum(t.values()).backward()
x = torch.randn(2, 8, 4, 4, device="cuda", requires_grad=True)
l, n = dr.template_features(x, dr.vertices_init)
(l.sum() + n.sum()).backward()
torch.cuda.synchronize()
print("case ok glue", flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -c 'case ok' gpurun_out/sanitize_$tool.log) cases, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
