"""Import the reference's own hot-path modules on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference exists (the build container); nothing that runs on the GPU box
may call this.  It is used by oracle/make_golden.py to produce tests/golden/*.npz and by the
optional `tests/test_oracle_vs_reference.py` cross-check.

Accommodations (none changes arithmetic):
  * `config` is the reference's config.py executed verbatim with DEVICE overridden to 'cpu';
  * the package __init__ files (which drag in kaolin, torch_geometric, trimesh, open3d) are
    bypassed by registering bare package objects whose __path__ points at the real directories;
  * modules/transform/rotate.py:34 `.to(DEVICE)` is a no-op on CPU and leaves a leaf tensor that is
    then written in place; the source is loaded with that one token replaced by `.clone()`;
  * `kaolin.rep.TriangleMesh` (absent) is stubbed by a container with from_obj/from_tensors/to,
    enough for modules/meshing/{sphere,cuboid,meshing}.py to run their own arithmetic;
  * modules/network/gcn.py (vertex-feature pooling) imports torch_geometric layers (absent) that
    its static pooling methods never touch: the four names are stubbed; `get_bound_of_images` allocates its result with
    a hard-coded `.cuda()` (gcn.py:87), loaded with that token removed.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("VPN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "loss", "chamfer_distance.py"))


class _StubTriangleMesh:
    """Minimal stand-in for kaolin.rep.TriangleMesh (v0.1 API surface used by meshing/*.py)."""

    def __init__(self, vertices, faces):
        self.vertices = vertices
        self.faces = faces

    @classmethod
    def from_obj(cls, path):
        from oracle.vpn_oracle import parse_obj
        with open(path) as fh:
            v, f = parse_obj(fh.read())
        return cls(torch.from_numpy(v.copy()), torch.from_numpy(f.copy()))

    @classmethod
    def from_tensors(cls, vertices, faces):
        return cls(vertices, faces)

    def to(self, device):
        self.vertices = self.vertices.to(device)
        self.faces = self.faces.to(device)
        return self


def _shadowed(name: str) -> bool:
    return name in ("config", "modules", "kaolin") or name.startswith("modules.") or name.startswith("kaolin.")


def _bare_package(name: str, path: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _load_patched(name: str, path: str, replace=None) -> types.ModuleType:
    with open(path) as fh:
        src = fh.read()
    if replace is not None:
        assert replace[0] in src, "patch anchor missing in " + path
        src = src.replace(replace[0], replace[1])
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__package__ = name.rpartition(".")[0]
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_reference() -> types.SimpleNamespace:
    """Returns a namespace with the reference's own functions/classes for the hot path."""
    if not available():
        raise RuntimeError("reference tree not found at " + REFERENCE_ROOT)
    if "_vpn_reference_ns" in sys.modules:
        return sys.modules["_vpn_reference_ns"]
    saved = {k: v for k, v in sys.modules.items() if _shadowed(k)}
    for k in saved:
        del sys.modules[k]
    root = REFERENCE_ROOT
    # config with DEVICE = 'cpu'
    cfg = _load_patched("config", os.path.join(root, "config.py"), ("DEVICE = 'cuda'", "DEVICE = 'cpu'"))
    # kaolin stub (only TriangleMesh is touched by the meshing files)
    kaolin = types.ModuleType("kaolin"); kaolin.__path__ = []
    rep = types.ModuleType("kaolin.rep"); rep.TriangleMesh = _StubTriangleMesh
    kaolin.rep = rep
    sys.modules["kaolin"], sys.modules["kaolin.rep"] = kaolin, rep
    # bare packages
    mroot = os.path.join(root, "modules")
    _bare_package("modules", mroot)
    tr_pkg = _bare_package("modules.transform", os.path.join(mroot, "transform"))
    _bare_package("modules.sampling", os.path.join(mroot, "sampling"))
    _bare_package("modules.loss", os.path.join(mroot, "loss"))
    ms_pkg = _bare_package("modules.meshing", os.path.join(mroot, "meshing"))
    rotate = _load_patched("modules.transform.rotate", os.path.join(mroot, "transform", "rotate.py"),
                           ("requires_grad=True).to(DEVICE)", "requires_grad=True).clone()"))
    translate = importlib.import_module("modules.transform.translate")
    transform = importlib.import_module("modules.transform.transform")
    for name in ("transform_points", "view_to_obj_points", "obj_to_view_points", "rotate_points_forward_x_axis"):
        setattr(tr_pkg, name, getattr(transform, name))
    tr_pkg.rotate_points = rotate.rotate_points
    sphere = importlib.import_module("modules.sampling.sphere")
    cuboid = importlib.import_module("modules.sampling.cuboid")
    sampling = importlib.import_module("modules.sampling.sampling")
    chamfer = importlib.import_module("modules.loss.chamfer_distance")
    vpdiv = importlib.import_module("modules.loss.vp_diverse")
    mesh_sphere = importlib.import_module("modules.meshing.sphere")
    mesh_cuboid = importlib.import_module("modules.meshing.cuboid")
    ms_pkg.cuboid, ms_pkg.sphere = mesh_cuboid, mesh_sphere
    meshing = importlib.import_module("modules.meshing.meshing")
    # GCN vertex-feature pooling: only the static methods are used; torch_geometric layers are never constructed
    tg = types.ModuleType("torch_geometric"); tg.__path__ = []
    tgnn = types.ModuleType("torch_geometric.nn")
    for nm in ("GCNConv", "TAGConv", "GraphUNet", "BatchNorm"):
        setattr(tgnn, nm, type(nm, (), {}))
    tg.nn = tgnn
    had_tg = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.nn")}
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = tg, tgnn
    _bare_package("modules.network", os.path.join(mroot, "network"))
    gcn = _load_patched("modules.network.gcn", os.path.join(mroot, "network", "gcn.py"),
                        ("torch.zeros((imgs.size(0), 4)).cuda()", "torch.zeros((imgs.size(0), 4))"))
    for k, v in had_tg.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    ns = types.SimpleNamespace(GCNModel=gcn.GCNModel,
        root=root, config=cfg, rotate=rotate, translate=translate, transform=transform,
        sphere=sphere, cuboid=cuboid, Sampling=sampling.Sampling,
        ChamferDistanceLoss=chamfer.ChamferDistanceLoss, VPDiverseLoss=vpdiv.VPDiverseLoss,
        mesh_sphere=mesh_sphere, mesh_cuboid=mesh_cuboid, Meshing=meshing.Meshing,
        TriangleMesh=_StubTriangleMesh)
    sys.modules["_vpn_reference_ns"] = ns
    # Every cross-module name the reference needs was bound at import time, so its entries can be
    # dropped from sys.modules again: the product's own drop-in package is also called `modules`
    # (and ships a `config`), and the two must be able to live in one test process.
    for k in [k for k in sys.modules if _shadowed(k)]:
        del sys.modules[k]
    sys.modules.update(saved)
    return ns


class forced_uniforms:
    """Context manager: torch.rand inside the reference samplers returns pre-drawn tensors, in call
    order (sphere.py:26-27 draws elev then azim, each (B,N,1); cuboid.py:66 draws one (B,N,3))."""

    def __init__(self, tensors):
        self.queue = list(tensors)

    def __enter__(self):
        self._orig = torch.rand

        def fake_rand(*size, **kw):
            shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
            t = self.queue.pop(0)
            assert tuple(t.shape) == shape, (tuple(t.shape), shape)
            return t.clone()

        torch.rand = fake_rand
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig
        return False
