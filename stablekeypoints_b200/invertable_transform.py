"""Mirror of the reference's ``unsupervised_keypoints/invertable_transform.py`` (:6-92) on the warp kernel."""
from __future__ import annotations

import math

import torch

from . import ops


def invert_theta(theta: torch.Tensor) -> torch.Tensor:
    """inverse([theta; 0 0 1])[:2] in closed form (invertable_transform.py:78-85), fp64 then fp32; [B,2,3]."""
    t = theta.detach().to("cpu", torch.float64)
    a, b, tx = t[:, 0, 0], t[:, 0, 1], t[:, 0, 2]
    c, d, ty = t[:, 1, 0], t[:, 1, 1], t[:, 1, 2]
    det = a * d - b * c
    inv = torch.stack([torch.stack([d / det, -b / det, (b * ty - d * tx) / det], -1),
                       torch.stack([-c / det, a / det, (c * tx - a * ty) / det], -1)], 1)
    return inv.to(torch.float32)


class RandomAffineWithInverse:
    def __init__(self, degrees=0, scale=(1.0, 1.0), translate=(0.0, 0.0)):
        self.degrees = degrees
        self.scale = scale
        self.translate = translate
        self.last_params = {"theta": torch.eye(2, 3).unsqueeze(0)}

    def create_affine_matrix(self, angle, scale, translations_percent):
        """invertable_transform.py:21-36."""
        a = math.radians(angle)
        theta = torch.tensor([[math.cos(a), math.sin(a), translations_percent[0]],
                              [-math.sin(a), math.cos(a), translations_percent[1]]], dtype=torch.float)
        theta[:, :2] = theta[:, :2] * scale
        return theta.unsqueeze(0)

    def sample_theta(self, batch: int) -> torch.Tensor:
        """invertable_transform.py:42-57: four host-side torch.rand(1) draws per image, in the reference's order."""
        theta = []
        for _ in range(batch):
            angle = torch.rand(1).item() * (2 * self.degrees) - self.degrees
            scale_factor = torch.rand(1).item() * (self.scale[1] - self.scale[0]) + self.scale[0]
            tr = (torch.rand(1).item() * (2 * self.translate[0]) - self.translate[0],
                  torch.rand(1).item() * (2 * self.translate[1]) - self.translate[1])
            theta.append(self.create_affine_matrix(angle, scale_factor, tr))
        return torch.cat(theta, dim=0)

    def __call__(self, img_tensor, theta=None):
        """invertable_transform.py:38-70; the warp itself runs on the GPU (the reference warps on the CPU)."""
        if theta is None:
            theta = self.sample_theta(img_tensor.shape[0])
        self.last_params = {"theta": theta}
        dev = img_tensor.device if img_tensor.is_cuda else torch.device("cuda", torch.cuda.current_device())
        return ops.affine_warp(img_tensor.to(dev, non_blocking=True), theta)

    def inverse(self, img_tensor):
        """invertable_transform.py:72-92."""
        return ops.affine_warp(img_tensor, invert_theta(self.last_params["theta"]))
