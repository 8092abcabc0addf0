"""Drop-in for modules/meshing.Meshing (reference: meshing/meshing.py:8-55, sphere.py, cuboid.py).

The reference parses the template OBJ from disk once per (primitive, sample) per iteration; here the
template lives on the device (vpn_b200.templates) and the vertex math is the fused pose kernel."""
import torch

from config import DEVICE
from vpn_b200 import ops, templates


class TriangleMesh:
    """The slice of kaolin.rep.TriangleMesh this path touches: .vertices (V,3) float, .faces (F,3) int64."""

    def __init__(self, vertices, faces, topology_key=None):
        self.vertices = vertices
        self.faces = faces
        self.topology_key = topology_key

    @classmethod
    def from_tensors(cls, vertices, faces):
        return cls(vertices, faces)

    @classmethod
    def from_obj(cls, path):
        v, f = templates.parse_obj(path)
        return cls(torch.from_numpy(v), torch.from_numpy(f).long(), topology_key=('obj', path))

    def sample(self, num_samples: int):
        """kaolin TriangleMesh.sample (train_sphere.py:76, dataset.py:165): (points (n,3), face_idx (n,) int64),
        area-weighted, drawn on the device; differentiable w.r.t. the vertices."""
        u = torch.rand((1, num_samples, 3), device=self.vertices.device)
        points, face_idx = ops.sample_mesh_surface(self.vertices[None], self.faces.to(torch.int32), u)
        return points[0], face_idx[0].long()

    @staticmethod
    def sample_batch(meshes: list, num_samples: int):
        """All meshes of a batch (shared topology) in two launches: what train_sphere.py:71-80 loops over."""
        verts = torch.stack([m.vertices for m in meshes])
        u = torch.rand((len(meshes), num_samples, 3), device=verts.device)
        return ops.sample_mesh_surface(verts, meshes[0].faces.to(torch.int32), u)[0]

    def to(self, device):
        self.vertices = self.vertices.to(device)
        self.faces = self.faces.to(device)
        return self

    def cuda(self):
        return self.to('cuda')


def _faces64(name, device):
    key = ('faces64', name, str(device))
    cache = _faces64.__dict__.setdefault('cache', {})
    if key not in cache:
        cache[key] = templates.template(name, device)[1].long()
    return cache[key]


class Meshing:
    def __init__(self):
        pass

    @classmethod
    def _meshing(cls, name, v, q, t):
        cls.check_parameters(v, q, t)
        tv, _ = templates.template(name, DEVICE)
        verts = ops.mesh_vertices(tv, v[:, None], q[:, None], t[:, None])       # (B, V, 3)
        faces = _faces64(name, DEVICE)
        return [TriangleMesh(verts[b], faces, topology_key=(name,)) for b in range(v.size(0))]

    @classmethod
    def cuboid_meshing(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor) -> list:
        return cls._meshing('cuboid', v, q, t)

    @classmethod
    def sphere_meshing(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor) -> list:
        return cls._meshing('sphere', v, q, t)

    @classmethod
    def cone_meshing(cls, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor) -> list:
        cls.check_parameters(v, q, t)
        pass                                                         # meshing.py:23-26: stub in the reference too

    @staticmethod
    def compose_meshes(meshes: list) -> TriangleMesh:
        """meshing.py:28-46: concatenate vertices; faces get the running vertex offset."""
        vertices, faces, vertices_num, keys = [], [], 0, []
        for mesh in meshes:
            vertices.append(mesh.vertices)
            faces.append(mesh.faces + vertices_num)
            vertices_num += mesh.vertices.size(0)
            keys.append(getattr(mesh, 'topology_key', None))
        key = None if any(k is None for k in keys) else ('composed',) + tuple(keys)
        return TriangleMesh(torch.cat(vertices), torch.cat(faces), topology_key=key).to(DEVICE)

    @staticmethod
    def check_parameters(v: torch.Tensor, q: torch.Tensor, t: torch.Tensor):
        assert v.size(0) == q.size(0) == t.size(0)
        B = v.size(0)
        assert v.size() == (B, 3)
        assert q.size() == (B, 4)
        assert t.size() == (B, 3)
