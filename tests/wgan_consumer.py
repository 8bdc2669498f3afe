"""TEST INFRASTRUCTURE: the WGAN-GP consumer of the rendered RGBA, restated from the reference for use on the
GPU box (where /root/reference does not exist): the critic's layer stack follows networks.py:87-144 (`Discriminator`: 1x1 and
3x3 convolutions without bias, LeakyReLU(0.2), stride-2 every other layer, spatial mean), the penalty follows
smr_utils.py:340-360 (`compute_gradient_penalty`: random interpolates, autograd.grad with create_graph -> (|g|_2 - 1)^2) and
the two steps follow trainer.py:391-438 (D: fake - real + gp; G: -D(fake)).  The alpha draws are seeded torch instead of
np.random so that CPU and CUDA arms see the same interpolates."""
import torch
import torch.nn as nn


def make_critic(nc=4, nf=16, seed=0, depth=3):
    """networks.py:87-131 with `depth` stride-2 stages (the reference uses 6 for 128^2 inputs; 3 keeps a 64^2 test quick)."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    layers = [nn.Conv2d(nc, nf, 1, 1, 0, bias=False), nn.LeakyReLU(0.2)]
    c = nf
    for i in range(depth):
        c2 = min(nf * (i + 2), nf * 4)
        layers += [nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.LeakyReLU(0.2), nn.Conv2d(c, c2, 3, 2, 1, bias=False), nn.LeakyReLU(0.2)]
        c = c2
    layers += [nn.Conv2d(c, nf * 2, 1, 1, 0, bias=False), nn.LeakyReLU(0.2), nn.Conv2d(nf * 2, 1, 1, 1, 0, bias=False)]
    net = nn.Sequential(*layers)
    for m in net:
        if isinstance(m, nn.Conv2d):
            nn.init.normal_(m.weight, 0.0, 0.05)
    torch.random.set_rng_state(st)

    class Critic(nn.Module):
        def __init__(self):
            super().__init__()
            self.main = net

        def forward(self, x):
            return self.main(x).mean([2, 3])
    return Critic()


def gradient_penalty(D, real, fake, alpha):
    """smr_utils.py:340-360 with the interpolation weights passed in."""
    interpolates = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    d = D(interpolates)
    g = torch.autograd.grad(outputs=d, inputs=interpolates, grad_outputs=torch.ones_like(d), create_graph=True,
                            retain_graph=True, only_inputs=True)[0]
    g = g.reshape(g.size(0), -1)
    return ((g.norm(2, dim=1) - 1) ** 2).mean()


def d_step(D, opt, real, fake, alpha, lambda_gan=1e-4, gan_reg=10.0):
    """trainer.py:391-418 (wgan branch, one fake stream): the double backward through the critic."""
    opt.zero_grad()
    out_r, out_f = D(real.detach()), D(fake.detach())
    loss = lambda_gan * out_f.mean() - lambda_gan * out_r.mean() + gan_reg * lambda_gan * gradient_penalty(D, real.detach(), fake.detach(), alpha)
    loss.backward()
    opt.step()
    return float(loss)


def g_loss(D, fake, lambda_gan=1e-4):
    """trainer.py:429-435 (wgan branch): the generator's adversarial term, whose gradient flows back INTO the render."""
    return lambda_gan * (-D(fake).mean())
