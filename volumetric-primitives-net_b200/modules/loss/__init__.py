"""Drop-in for modules/loss (reference: loss/chamfer_distance.py, vp_diverse.py, silhouette.py, emd/emd_module.py)."""
import torch
import torch.nn as nn

from config import CD_W1, CD_W2, SILHOUETTE_LOSS_FUNC, VP_NUM
from vpn_b200 import ops


class ChamferDistanceLoss(nn.Module):
    """chamfer_distance.py:6-35."""

    def __init__(self):
        super().__init__()

    def forward(self, points1: torch.Tensor, points2: torch.Tensor, each_batch=False, w1=CD_W1, w2=CD_W2) -> torch.Tensor:
        self.check_parameters(points1)
        self.check_parameters(points2)
        return ops.chamfer_distance(points1, points2, each_batch=each_batch, w1=w1, w2=w2)

    @staticmethod
    def check_parameters(points: torch.Tensor):
        assert points.ndimension() == 3  # (B, N, 3)
        assert points.size(-1) == 3


class VPDiverseLoss(nn.Module):
    """vp_diverse.py:7-24."""

    def __init__(self):
        super().__init__()
        self.cd_loss_func = ChamferDistanceLoss()

    def forward(self, translates: list, gt_points: torch.Tensor) -> torch.Tensor:
        self.check_parameters(translates)
        vp_center_points = torch.cat([t[:, None, :] for t in translates], 1)
        return self.cd_loss_func(vp_center_points, gt_points, w1=0.5, w2=1.0)

    @staticmethod
    def check_parameters(translates):
        assert isinstance(translates, list)
        assert len(translates) == VP_NUM


class SilhouetteLoss(nn.Module):
    """silhouette.py:8-23.  Meshes that share one topology (the normal case: composed template
    primitives) are rendered as ONE batched launch instead of the reference's per-sample loop."""

    def __init__(self):
        super().__init__()
        self.loss_func = nn.L1Loss() if SILHOUETTE_LOSS_FUNC == 'L1' else nn.MSELoss()

    def forward(self, predict_meshes: list, gt_silhouettes: torch.Tensor,
                dists: torch.Tensor, elevs: torch.Tensor, azims: torch.Tensor) -> torch.Tensor:
        from modules.render import render_alpha_batch
        h, w = gt_silhouettes.shape[-2:]
        predict_silhouettes = render_alpha_batch(predict_meshes, dists, elevs, azims, h, w)   # (B,1,H,W)
        return self.loss_func(predict_silhouettes, gt_silhouettes)


class EarthMoverDistanceLoss(nn.Module):
    """emd/emd_module.py:72-79: forward(input1, input2, eps, iters) -> (dist (B,n), assignment (B,n) int32).
    Same preconditions as the reference (emd_module.py:37-40)."""

    def __init__(self):
        super().__init__()

    def forward(self, input1, input2, eps, iters):
        batchsize, n, _ = input1.size()
        _, m, _ = input2.size()
        assert n == m
        assert input1.size()[0] == input2.size()[0]
        assert n % 1024 == 0
        assert batchsize <= 512
        return ops.emd_auction(input1.contiguous().float(), input2.contiguous().float(), eps, iters)
