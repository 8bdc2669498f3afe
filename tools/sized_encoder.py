"""BENCH INFRASTRUCTURE: an attribute encoder with the reference's SIZE and SHAPE, for the trainer-step benches (cfg-3 / cfg-4).

The reference's `AttributeEncoder` (networks.py:532-647 over network/model_res.py:84-612) is five CNNs -- camera 6.50 M, shape
9.55 M, light 1.14 M, texture 16.49 M, background 0.03 M parameters = 33.7 M (counted by running the real classes) --
whose backbones need `timm` + files that are not on the GPU box.  What a trainer STEP needs from them is (a) the same parameter
volume for the gradient all-reduce (135 MB fp32), (b) BatchNorm per rank, (c) the same order of convolution work and (d) the
attribute dict with the reference's keys / ranges / post-processing.  This module provides exactly that with plain ResNet-style
trunks (tensor cores through cuDNN: bf16 autocast + channels_last) -- it is NOT the reference's architecture and no accuracy
claim hangs on it.  Two of the product's own kernels sit where the reference has the corresponding ops:
  * ShapeEncoder's template conditioning (model_res.py:317-325)  -> DiffRender.template_features
  * TextureEncoder's bicubic flow sampling + flip-concat (model_res.py:598-610) -> DiffRender.texture_flow
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.sk = None if (stride == 1 and cin == cout) else nn.Sequential(nn.Conv2d(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)), inplace=True)
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.sk is None else self.sk(x)), inplace=True)


class Trunk(nn.Module):
    """ResNet-18-shaped: stem /2, four stages of two blocks, /32 overall; returns the four stage outputs."""

    def __init__(self, cin, widths):
        super().__init__()
        w0 = widths[0]
        self.stem = nn.Sequential(nn.Conv2d(cin, w0, 7, 2, 3, bias=False), nn.BatchNorm2d(w0), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1))
        stages, c = [], w0
        for i, w in enumerate(widths):
            stages.append(nn.Sequential(Block(c, w, 1 if i == 0 else 2), Block(w, w, 1)))
            c = w
        self.stages = nn.ModuleList(stages)

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for s in self.stages:
            x = s(x)
            outs.append(x)
        return outs


class SizedEncoder(nn.Module):
    def __init__(self, dr, H, W, nf_shape=288, amp=True):
        super().__init__()
        self.dr, self.H, self.W, self.amp = dr, H, W, amp      # amp=False: fp32 convolutions (equality tests)
        V = dr.num_vertices
        self.V = V
        self.register_buffer("vertices_init", dr.vertices_init.clone()[None])
        self.register_buffer("light_mean", torch.tensor([3.0] + [0.0] * 8))
        self.register_buffer("light_scale", torch.tensor([0.5] + [0.1] * 8))
        # camera: 6.5 M
        self.cam_trunk = Trunk(4, [48, 96, 192, 400])
        self.cam_head = nn.Linear(400, 6)
        # shape: 9.55 M  (trunk -> (B, nf_shape, H/32, W/32) -> template_features -> per-vertex MLP)
        self.shape_trunk = Trunk(4, [56, 112, 224, 448])
        self.shape_proj = nn.Conv2d(448, nf_shape, 1, bias=False)
        self.shape_mlp = nn.Sequential(nn.Conv1d(2 * nf_shape + 3, 768, 1), nn.ReLU(inplace=True), nn.Conv1d(768, 384, 1), nn.ReLU(inplace=True),
                                       nn.Conv1d(384, 3, 1))
        # light: 1.14 M
        self.light_trunk = Trunk(4, [20, 40, 80, 160])
        self.light_head = nn.Linear(160, 9)
        # texture: 16.5 M  (encoder + top-down decoder to a full-resolution 2-channel flow)
        tw = [64, 128, 256, 528]
        self.tex_trunk = Trunk(4, tw)
        self.tex_up = nn.ModuleList([nn.Sequential(nn.Conv2d(tw[i] + (tw[i + 1] if i < 3 else 0), tw[i], 3, 1, 1, bias=False), nn.BatchNorm2d(tw[i]),
                                                   nn.ReLU(inplace=True)) for i in range(4)])
        self.tex_mid = nn.Sequential(nn.Conv2d(64, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                     nn.Conv2d(64, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        self.tex_out = nn.Conv2d(64, 2, 3, 1, 1)
        # background: 0.03 M
        self.bg = nn.Sequential(nn.Conv2d(4, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(32, 64, 3, 1, 1), nn.ReLU(inplace=True),
                                nn.Conv2d(64, 3, 3, 1, 1))

    def param_counts(self):
        n = lambda *ms: sum(p.numel() for m in ms for p in m.parameters())      # noqa: E731
        return {"camera": n(self.cam_trunk, self.cam_head), "shape": n(self.shape_trunk, self.shape_proj, self.shape_mlp),
                "light": n(self.light_trunk, self.light_head), "texture": n(self.tex_trunk, self.tex_up, self.tex_mid, self.tex_out),
                "bg": n(self.bg), "total": n(self)}

    def forward(self, img):
        B = img.shape[0]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            x = img.contiguous(memory_format=torch.channels_last)
            cam = torch.tanh(self.cam_head(self.cam_trunk(x)[-1].mean(dim=(2, 3)).float()))
            lights = self.light_mean + self.light_scale * torch.tanh(self.light_head(self.light_trunk(x)[-1].mean(dim=(2, 3)).float()))
            feat = self.shape_proj(self.shape_trunk(x)[-1])                     # (B, 288, H/32, W/32)
            ts = self.tex_trunk(x)
            y = None
            for i in (3, 2, 1, 0):                                              # top-down decoder
                z = ts[i] if y is None else torch.cat([ts[i], F.interpolate(y, size=ts[i].shape[2:], mode='nearest')], 1)
                y = self.tex_up[i](z)
            y = self.tex_mid(F.interpolate(y, size=(self.H, self.W), mode='bilinear', align_corners=False))
            flow = torch.tanh(self.tex_out(y)).float()                          # (B, 2, H, W) in [-1, 1]
            bg = torch.sigmoid(self.bg(x)).float().contiguous()
        # the render path and the two glue kernels run in fp32, outside autocast (trainer.py:271-276)
        local, ndiff = self.dr.template_features(feat.float().contiguous(), self.vertices_init)       # (B,288,V,1) each
        tmpl = self.vertices_init.transpose(1, 2).expand(B, 3, self.V)
        pv = torch.cat([local.squeeze(-1), ndiff.squeeze(-1), tmpl], dim=1)      # (B, 579, V)
        delta = 0.05 * torch.tanh(self.shape_mlp(pv)).transpose(1, 2)            # (B, V, 3)
        delta = delta - delta.mean(dim=1, keepdim=True)
        textures = self.dr.texture_flow(img[:, :3].contiguous(), flow.contiguous(), concat=True)      # (B, 3, 2H, W)
        return {
            'azimuths': 180.0 * cam[:, 0], 'elevations': 15.0 * cam[:, 1], 'distances': 4.0 + 2.0 * cam[:, 2],
            'biases': 0.5 * cam[:, 3:5], 'vertices': self.vertices_init + delta, 'delta_vertices': delta,
            'textures': textures, 'lights': lights, 'bg': bg,
        }
