"""Drop-in for the hot-path part of the reference's `modules` package (modules/__init__.py:1-8).

Same names, call signatures and AssertionError behaviour as the reference for transform, sampling,
loss (Chamfer / VP-diverse / silhouette / EMD auction), render (VertexRenderer), meshing and the GCN vertex-feature
pooling (modules.pooling; gcn.py:84-164); everything
underneath runs as vpn_b200 CUDA kernels.  Networks, datasets, augmentation and visualisation are out of
scope and stay with the reference.
"""
from .meshing import Meshing
from .sampling import Sampling
from .loss import ChamferDistanceLoss, EarthMoverDistanceLoss, SilhouetteLoss, VPDiverseLoss
from .render import VertexRenderer
from .pooling import GCNFeaturePooling
from .transform import (transform_points, view_to_obj_points, obj_to_view_points, rotate_points,
                        rotate_points_forward_x_axis, translate_points)
