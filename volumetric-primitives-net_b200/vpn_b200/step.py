"""One primitive-loss step: the per-iteration primitive assembly + loss of train.py:243-258 (EMD excluded),
batched over all K primitives and all B samples, every stage a vpn_b200 kernel.

    points   = sample + pose of K primitives            (train.py:105-120 -> one launch)
    view_cd  = Chamfer(points, view_center_points)       (train.py:152-163)
    obj_cd   = Chamfer(view_to_obj(points), canonical)   (only when L_CAN_CD != 0; the reference computes
                                                          it even when its weight is 0, train.py:160-161)
    vp_div   = Chamfer(centres, targets, w1=.5, w2=1)    (train.py:179-185, vp_diverse.py:12-18)
    sil      = L1/MSE(soft_silhouette(mesh), gt)         (train.py:123-149,166-176; silhouette.py:13-23)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops, templates


@dataclass
class PrimitiveLossConfig:
    kind: str = "sphere"              # 'sphere' | 'cuboid'   (config.py:32-35 SPHERE_NUM / CUBOID_NUM)
    l_view_cd: float = 1.0            # config.py:14-17
    l_can_cd: float = 0.0
    l_sil: float = 0.0
    l_vp_div: float = 0.1
    cd_w1: float = 1.0                # config.py:12-13
    cd_w2: float = 1.0
    silhouette_loss: str = "L1"       # config.py:27
    img_size: int = 128               # config.py:46
    chamfer_impl: int = ops.CHAMFER_AUTO
    vertex_chamfer: bool = False      # train_gcn.py:127-130: the Chamfer term scores the composed mesh vertices, not samples


class PrimitiveLoss:
    def __init__(self, config: Optional[PrimitiveLossConfig] = None):
        self.cfg = config or PrimitiveLossConfig()
        assert self.cfg.kind in ("sphere", "cuboid")

    def composed_faces(self, k: int, device) -> torch.Tensor:
        """Faces of K composed template meshes (meshing.py:28-46 running vertex offset), int32 (K*F,3)."""
        key = (k, str(device))
        cache = self.__dict__.setdefault("_faces", {})
        if key not in cache:
            tv, tf = templates.template(self.cfg.kind, device)
            nv = tv.shape[0]
            cache[key] = torch.cat([tf + i * nv for i in range(k)], dim=0).contiguous()
        return cache[key]

    def default_cameras(self, b: int, device):
        """IS_VIEW_CENTER cameras: dist = 1, elev = azim = 0 for every sample (train.py:172-174); built once per batch size."""
        key = ("cam", b, str(device))
        cache = self.__dict__.setdefault("_cams", {})
        if key not in cache:
            cache[key] = ops.look_at_cameras(torch.zeros(b, device=device), torch.zeros(b, device=device),
                                             torch.ones(b, device=device))
        return cache[key]

    def __call__(self, v: torch.Tensor, q: torch.Tensor, t: torch.Tensor, uniforms: torch.Tensor,
                 view_center_points: torch.Tensor, silhouettes: Optional[torch.Tensor] = None,
                 canonical_points: Optional[torch.Tensor] = None, dists=None, elevs=None, azims=None,
                 angles=None) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        out: Dict[str, torch.Tensor] = {}
        b, k = q.shape[:2]
        verts = None
        if cfg.vertex_chamfer:
            tv, _ = templates.template(cfg.kind, v.device)
            points = verts = ops.mesh_vertices(tv, v, q, t)
        else:
            points = ops.sample_primitives(cfg.kind, v, q, t, uniforms)
        out["points"] = points
        total = None

        def add(name, val):
            nonlocal total
            out[name] = val
            total = val if total is None else total + val

        if cfg.l_view_cd:
            add("view_cd", ops.chamfer_distance(points, view_center_points, w1=cfg.cd_w1, w2=cfg.cd_w2,
                                                impl=cfg.chamfer_impl) * cfg.l_view_cd)
        if cfg.l_can_cd:
            can = ops.view_to_obj_points(points, dists, elevs, azims, angles)
            add("obj_cd", ops.chamfer_distance(can, canonical_points, w1=cfg.cd_w1, w2=cfg.cd_w2,
                                               impl=cfg.chamfer_impl) * cfg.l_can_cd)
        if cfg.l_vp_div:
            add("vp_div", ops.chamfer_distance(t.contiguous(), view_center_points, w1=0.5, w2=1.0,
                                               impl=cfg.chamfer_impl) * cfg.l_vp_div)
        if cfg.l_sil:
            if verts is None:
                tv, _ = templates.template(cfg.kind, v.device)
                verts = ops.mesh_vertices(tv, v, q, t)
            faces = self.composed_faces(k, v.device)
            h, w = silhouettes.shape[-2:]
            if azims is None and elevs is None and dists is None:
                rot, pos = self.default_cameras(b, v.device)
            else:
                one = torch.ones(b, device=v.device)
                zero = torch.zeros(b, device=v.device)
                rot, pos = ops.look_at_cameras(zero if azims is None else azims, zero if elevs is None else elevs,
                                               one if dists is None else dists)
            alpha, _, _ = ops.soft_silhouette(verts, faces, rot, pos, h, w)
            diff = alpha[:, None] - silhouettes
            sil = diff.abs().mean() if cfg.silhouette_loss == "L1" else (diff * diff).mean()
            out["alpha"] = alpha
            add("sil", sil * cfg.l_sil)
        out["total"] = total
        return out
