"""Drop-in for modules/render.VertexRenderer (reference: render/vertex_renderer.py:1-36).

The reference keeps one module-global kaolin DIBRenderer(128, 128) whose camera is mutated per call and
reads the camera scalars back with .item() (three host syncs per sample).  Here the camera is built on
the device from the tensors and the renderer is stateless."""
import torch

from config import DEVICE, IMG_SIZE
from vpn_b200 import ops


def _as_tensor(x, device):
    if isinstance(x, torch.Tensor):
        return x.reshape(-1).to(device=device, dtype=torch.float32)
    return torch.tensor([float(x)], dtype=torch.float32, device=device)


def _same_topology(meshes) -> bool:
    f0 = meshes[0].faces
    for m in meshes[1:]:
        if m.faces is f0:
            continue
        k0, k1 = getattr(meshes[0], 'topology_key', None), getattr(m, 'topology_key', None)
        if k0 is None or k0 != k1 or m.vertices.shape != meshes[0].vertices.shape:
            return False
    return True


def render_alpha_batch(meshes: list, dists, elevs, azims, height: int, width: int) -> torch.Tensor:
    """Soft alpha of every mesh: (B,1,H,W)."""
    dev = meshes[0].vertices.device
    b = len(meshes)
    d, e, a = _as_tensor(dists, dev), _as_tensor(elevs, dev), _as_tensor(azims, dev)
    rot, pos = ops.look_at_cameras(a, e, d)
    if _same_topology(meshes):
        verts = torch.stack([m.vertices for m in meshes])
        alpha, _, _ = ops.soft_silhouette(verts, meshes[0].faces.to(torch.int32), rot, pos, height, width)
        return alpha[:, None]
    outs = []
    for i in range(b):
        alpha, _, _ = ops.soft_silhouette(meshes[i].vertices[None], meshes[i].faces.to(torch.int32),
                                          rot[i:i + 1], pos[i:i + 1], height, width)
        outs.append(alpha[:, None])
    return torch.cat(outs)


class VertexRenderer:
    def __init__(self):
        pass

    @classmethod
    def render(cls, mesh, dist, elev, azim, colors=None):
        """vertex_renderer.py:15-26: returns (rgb (1,H,W,3), alpha (1,H,W,1), face_normals (1,F,3)).
        Vertex colours default to ones; the colour image is then the hard coverage mask."""
        dev = torch.device(DEVICE)
        vertices = mesh.vertices.to(dev)[None]
        faces = mesh.faces.to(dev).to(torch.int32)
        rot, pos = ops.look_at_cameras(_as_tensor(azim, dev), _as_tensor(elev, dev), _as_tensor(dist, dev))
        alpha, covered, normals = ops.soft_silhouette(vertices, faces, rot, pos, IMG_SIZE, IMG_SIZE, True)
        if colors is not None:
            raise NotImplementedError("per-vertex colours are not on the silhouette-loss path "
                                      "(silhouette.py:17 keeps only the alpha channel)")
        rgb = covered.to(torch.float32)[..., None].expand(-1, -1, -1, 3).contiguous()
        return rgb, alpha[..., None], normals
