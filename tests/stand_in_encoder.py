"""TEST INFRASTRUCTURE: a small stand-in for the reference's `AttributeEncoder` (networks.py:533-647).

The real encoders (network/model_res.py) need `timm` backbones and pretrained weights that are not in this image; what the
render path sees of them is the attribute dict (`networks.py:635-647`).  This module produces that dict from a 4-channel
image with the same keys, shapes, value ranges and post-processing as the reference heads:
    azimuths / elevations / distances / biases   CameraEncoder ranges (train.py:123-127)
    delta_vertices = 0.05 * tanh(.), mean-centred; vertices = vertices_init + delta   (model_res.py:333-337, networks.py:622)
    textures = cat([t, t.flip(2)], 2)                                                  (model_res.py:609-610)
    lights = [3,0,..] + [.5,.1,..] * tanh(.)                                           (model_res.py:392-395)
    bg                                                                                 (BackgroundEncoder)
No normalisation layers (per-rank BatchNorm statistics would make a 2-rank run differ from the single-process one by design)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class StandInEncoder(nn.Module):
    def __init__(self, vertices_init, H, W, tex_mirror=False):
        super().__init__()
        V = vertices_init.shape[0]
        self.H, self.W, self.V = H, W, V
        self.tex_mirror = tex_mirror                 # True: hand the renderer the un-concatenated half (render(_tex_mirror=True))
        self.register_buffer("vertices_init", vertices_init.clone()[None])
        self.c1 = nn.Conv2d(4, 16, 5, 2, 2)
        self.c2 = nn.Conv2d(16, 32, 5, 2, 2)
        self.c3 = nn.Conv2d(32, 64, 3, 2, 1)
        self.cam = nn.Linear(64, 5)
        self.shape = nn.Linear(64, V * 3)
        self.light = nn.Linear(64, 9)
        self.tex = nn.Conv2d(32, 3, 3, 1, 1)
        self.bg = nn.Conv2d(4, 3, 3, 1, 1)
        self.register_buffer("light_mean", torch.tensor([3.0] + [0.0] * 8))
        self.register_buffer("light_scale", torch.tensor([0.5] + [0.1] * 8))

    def forward(self, img):
        B = img.shape[0]
        x1 = F.leaky_relu(self.c1(img), 0.2)
        x2 = F.leaky_relu(self.c2(x1), 0.2)
        x3 = F.leaky_relu(self.c3(x2), 0.2)
        g = x3.mean(dim=(2, 3))
        cam = torch.tanh(self.cam(g))
        delta = 0.05 * torch.tanh(self.shape(g)).view(B, self.V, 3)
        delta = delta - delta.mean(dim=1, keepdim=True)
        t = torch.sigmoid(F.interpolate(self.tex(x2), size=(self.H, self.W), mode='bilinear', align_corners=False))
        textures = t if self.tex_mirror else torch.cat([t, t.flip([2])], dim=2)
        return {
            'azimuths': 180.0 * cam[:, 0],
            'elevations': 15.0 + 15.0 * cam[:, 1],
            'distances': 4.0 + 2.0 * cam[:, 2],
            'biases': 0.3 * cam[:, 3:5],
            'vertices': self.vertices_init + delta,
            'delta_vertices': delta,
            'textures': textures,
            'lights': self.light_mean + self.light_scale * torch.tanh(self.light(g)),
            'img_feats': None,
            'bg': torch.sigmoid(self.bg(img)),
        }


def make_encoder(vertices_init, H, W, seed, tex_mirror=False):
    """Seeded weights, created on CPU so that every process / device starts from identical parameters."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    enc = StandInEncoder(vertices_init.cpu(), H, W, tex_mirror=tex_mirror)
    torch.random.set_rng_state(gen_state)
    return enc


def make_images(B, H, W, seed):
    """Synthetic CUB-shape batch: RGB U[0,1] with a centred disc mask in the alpha channel."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    r = 0.30 + 0.1 * torch.rand(B, 1, 1, generator=g)
    disc = (((xx - W / 2 + 0.5) / W) ** 2 + ((yy - H / 2 + 0.5) / H) ** 2)[None] < r ** 2
    return torch.cat([rgb, disc.float().unsqueeze(1)], dim=1)


def trainer_step_loss(dr, enc, images, contour=0.1, lambda_reg=0.1, render=None, recon=None, regs=None):
    """The data + regularisation part of one trainer.py iteration (trainer.py:271-276, 441, 54-72, 505-509):
    encode -> render -> recon_data + mesh regularisers.  `render` / `recon` default to the DiffRender under test; the CPU
    oracle arm passes its own (same signature)."""
    Ae = enc(images)
    if render is None:
        extra = {'_tex_mirror': True} if getattr(enc, 'tex_mirror', False) else {}
        Xer, Ae = dr.render(no_mask=True, **extra, **Ae)
    else:
        Xer, Ae = render(Ae)
    loss_data = (recon or (lambda p, g: dr.recon_data(p, g, no_mask=True, contour=contour)))(Xer, images)
    R = regs if regs is not None else dr          # the CPU oracle arm passes tests/reg_torch.TorchRegularisers (the product has no CPU path)
    reg = R.calc_reg_loss(Ae) + 0.1 * R.calc_reg_deform(Ae['delta_vertices']) + 0.01 * R.calc_reg_depth(Ae['vertices']) \
        + 0.1 * R.calc_reg_edge(Ae['vertices'])
    return loss_data + lambda_reg * reg, Xer
