"""Drop-in for modules/sampling (reference: sampling/sampling.py:12-52)."""
import torch

from config import DEVICE
from vpn_b200 import ops


class Sampling:
    """Surface samples of one posed primitive per batch item.  The uniforms are drawn with torch.rand
    on DEVICE in the reference's own call order and shapes (sphere.py:26-27: two (B,N,1) draws, elev
    first; cuboid.py:66: one (B,N,3) draw), so a seeded run consumes the same random stream."""

    def __init__(self):
        pass

    @classmethod
    def cuboid_sampling(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor, num_points: int = 1000):
        cls.check_parameters(v, q, t)
        assert type(num_points) == int and num_points > 0          # cuboid.py:25-27
        B = v.size(0)
        u = torch.rand((B, num_points, 3), dtype=torch.float, device=DEVICE)
        return ops.sample_primitives('cuboid', v[:, None], q[:, None], t[:, None], u[:, None])

    @classmethod
    def sphere_sampling(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor, num_points: int = 1000):
        cls.check_parameters(v, q, t)
        assert type(num_points) == int and num_points > 0          # sphere.py:17-19
        B = v.size(0)
        u_elev = torch.rand((B, num_points, 1), device=DEVICE)
        u_azim = torch.rand((B, num_points, 1), device=DEVICE)
        u = torch.cat([u_elev, u_azim], dim=2)
        return ops.sample_primitives('sphere', v[:, None], q[:, None], t[:, None], u[:, None])

    @classmethod
    def cone_sampling(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor, num_points: int = 1000):
        pass                                                         # sampling.py:40-46: stub in the reference too

    @staticmethod
    def check_parameters(v, q, t):
        B = v.size(0)
        assert v.size() == (B, 3)
        assert q.size() == (B, 4)
        assert t.size() == (B, 3)
