"""Camera helpers with the reference's names and semantics (smr_utils.py:257-311).

Inside `DiffRender.render` these are fused into the CUDA vertex stage
(csrc/mm_vertex.cu); the torch versions below exist because reference scripts
also call them directly (e.g. show_camera.py) and because tests compare the fused
stage against them.  Plain tensor math, any device.
"""
import math

import torch


def camera_position_from_spherical_angles(dist, elev, azim, degrees=True):
    """(N,3) camera position from distance / elevation / azimuth (smr_utils.py:257-281)."""
    if degrees:
        elev = math.pi / 180.0 * elev
        azim = math.pi / 180.0 * azim
    cos_e = torch.cos(elev)
    pos = torch.stack([dist * cos_e * torch.sin(azim), dist * torch.sin(elev), dist * cos_e * torch.cos(azim)], dim=1)
    return pos.reshape(-1, 3)


def generate_transformation_matrix(camera_position, look_at, camera_up_direction):
    """(N,4,3) world->camera transform, P_cam = [P_world, 1] @ T (smr_utils.py:284-311)."""
    z_axis = camera_position - look_at
    z_axis = z_axis / z_axis.norm(dim=1, keepdim=True)
    x_axis = torch.cross(camera_up_direction, z_axis, dim=1)
    x_axis = x_axis / x_axis.norm(dim=1, keepdim=True)
    y_axis = torch.cross(z_axis, x_axis, dim=1)
    rot = torch.stack([x_axis, y_axis, z_axis], dim=2)
    trans = -camera_position.unsqueeze(1) @ rot
    return torch.cat([rot, trans], dim=1)
