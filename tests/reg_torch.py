"""TEST INFRASTRUCTURE: plain-torch statement of the reference's mesh regularisers (networks.py:392-491).

The product package has no PyTorch / CPU backend: `DiffRender.calc_reg_*` / `recon_flip` only run the fused CUDA kernel
(csrc/mm_meshreg.cu).  This module is the checker for that kernel.  It is itself pinned against the UNMODIFIED reference
(tests/test_host_setup.py: values + gradients from tests/golden/reg_*.npz, and a live comparison when /root/reference exists).
"""
import torch


class TorchRegularisers(object):
    """Same method names / arguments as the reference's DiffRender, evaluated with torch ops on the tensors' device."""

    def __init__(self, dr):
        self.dr = dr

    def recon_flip(self, att, L1):
        dr = self.dr
        Na = att['delta_vertices']
        idx = dr.flip_index.to(Na.device)
        Nf = Na.index_select(1, idx)
        Nf = Nf * Nf.new_tensor([1.0, 1.0, -1.0])
        diff = Na - Nf
        loss_norm = torch.abs(diff) if L1 else diff.norm(dim=2)
        mask_a = torch.relu(torch.sign(Na[:, :, 2]) * dr.sign_init.to(Na.device))
        mask_f = mask_a.index_select(1, idx)
        if L1:
            # reference broadcasting: (B,V,3) * (B,V) is only valid when V == 3; its intent per vertex
            return torch.mean(loss_norm * mask_f.unsqueeze(-1))
        return torch.mean(loss_norm * mask_f)

    def calc_reg_loss(self, att):
        dr = self.dr
        delta = att['delta_vertices']
        dev = delta.device
        lap = dr.vertices_laplacian_matrix.to(dev)
        e2f = dr.edge2faces.to(dev)
        fn = att['face_normals']
        nb_vertices = delta.shape[1]
        loss_laplacian = torch.mean(torch.matmul(lap, delta) ** 2) * nb_vertices * 3
        cos = torch.sum(fn[:, e2f[:, 0]] * fn[:, e2f[:, 1]], dim=2)
        loss_flat = torch.mean((cos - 1) ** 2) * e2f.shape[0]
        return dr.lambda_lpl * loss_laplacian + dr.lambda_flat * loss_flat

    def calc_reg_edge(self, pred):
        e = self.dr.edges.to(pred.device)
        length = torch.norm(pred[:, e[:, 0]] - pred[:, e[:, 1]], p=2, dim=2)
        bias = length - torch.mean(length, dim=1, keepdim=True)
        return 0.1 * torch.mean(torch.norm(bias, p=2, dim=1))

    def calc_reg_depth(self, pred):
        return torch.mean(pred[:, :, 2] ** 2)

    def _depth_weighted(self, pred, weight, eps):
        s = self.dr.sign_init.to(pred.device)
        z = pred[:, :, 2]
        return torch.mean((s >= 0) * (z - eps) ** 2 * weight + (s < 0) * (z + eps) ** 2 * weight)

    def calc_reg_depthR(self, pred, temp=2, eps=0.001):
        x = pred[:, :, 0].detach()
        y = pred[:, :, 1].detach()
        return self._depth_weighted(pred, torch.exp(temp * (x ** 2 + (y / self.dr.ratio) ** 2)), eps)

    def calc_reg_depthC(self, pred, eps=0.001):
        x = pred[:, :, 0].detach()
        y = pred[:, :, 1].detach()
        return self._depth_weighted(pred, x ** 2 + (y / self.dr.ratio) ** 2, eps)

    def calc_reg_deform(self, pred):
        b = pred.shape[0]
        return torch.mean(torch.norm(pred.reshape(-1, pred.size(2)), p=2, dim=1).reshape(b, -1))
