"""GCNModel's feature-pooling methods (modules/network/gcn.py:84-164) on vpn_b200 kernels: same names, arguments,
assertions and output layout."""
import torch

from vpn_b200 import ops


class GCNFeaturePooling:
    @classmethod
    def get_local_features(cls, vertices: torch.Tensor, rgbs: torch.Tensor, perceptual_features: list):
        # gcn.py:84-88
        bounds = cls.get_bound_of_images(rgbs)
        return cls.perceptual_feature_pooling(perceptual_features, vertices, bounds)

    @staticmethod
    def get_bound_of_images(imgs: torch.Tensor):
        assert imgs.ndimension() == 4  # (B, C, H, W)            gcn.py:91
        return ops.image_bounds(imgs, 0.03)  # (B, 4)

    @staticmethod
    def perceptual_feature_pooling(perceptual_features: list, points: torch.Tensor, bounds: torch.Tensor):
        assert points.ndimension() == 3.  # (B, N, 3)             gcn.py:137-138
        assert bounds.ndimension() == 2.  # (B, 4)
        return ops.perceptual_feature_pooling(perceptual_features, points, bounds)  # (B, N, C)
