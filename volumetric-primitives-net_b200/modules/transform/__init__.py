"""Drop-in for modules/transform (reference: transform/__init__.py:1, rotate.py, translate.py, transform.py)."""
import torch

from vpn_b200 import ops


def _check_points(points):
    assert points.ndimension() == 3  # (B, N, 3)          rotate.py:49-51
    assert points.size(-1) == 3


def rotate_points(points: torch.Tensor, quaternions: torch.Tensor):
    """rotate.py:7-25.  quaternions (B,4) = (axis, turn fraction in [0,1] -> [0, 2pi])."""
    _check_points(points)
    assert quaternions.ndimension() == 2 and quaternions.size(-1) == 4   # rotate.py:54-56
    return ops.transform_points(points, quaternions, None)


def translate_points(points: torch.Tensor, translations: torch.Tensor):
    """translate.py:4-8 (pure broadcast add; no kernel of its own is worth a launch)."""
    _check_points(points)
    assert translations.ndimension() == 2 and translations.size(-1) == 3
    return points + translations.unsqueeze(1).expand_as(points)


def transform_points(points: torch.Tensor, q: torch.Tensor, t: torch.Tensor):
    """transform.py:6-18."""
    assert points.ndimension() == 3 and points.size(-1) == 3
    B = points.size(0)
    assert q.size() == (B, 4)
    assert t.size() == (B, 3)
    return ops.transform_points(points, q, t)


def view_to_obj_points(points, dists, elevs, azims, angles):
    """transform.py:21-47."""
    assert points.ndimension() == 3
    assert dists.ndimension() == elevs.ndimension() == azims.ndimension() == 1
    return ops.view_to_obj_points(points, dists, elevs, azims, angles)


def obj_to_view_points(points, dists, elevs, azims):
    """transform.py:50-73."""
    assert points.ndimension() == 3
    assert dists.ndimension() == elevs.ndimension() == azims.ndimension() == 1
    return ops.obj_to_view_points(points, dists, elevs, azims)


def rotate_points_forward_x_axis(points: torch.Tensor, angles: torch.Tensor):
    """transform.py:76-94: rotate about +x by angles (degrees, [0, 360])."""
    assert points.ndimension() == 3
    assert angles.ndimension() == 1
    B = points.size(0)
    x = torch.tensor([[1.0, 0.0, 0.0]], device=points.device).expand(B, 3)
    q = torch.cat([x, angles.view(-1, 1) / 360], dim=1)
    return ops.transform_points(points, q, None)
