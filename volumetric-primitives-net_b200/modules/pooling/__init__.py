"""Drop-in for the vertex-feature pooling of the reference's GCN (modules/network/gcn.py:84-164).

The networks themselves stay with the reference (out of scope), so this package is NOT called `modules.network`.  The
three pooling methods of `GCNModel` are static / class methods that touch no parameters; the reference adopts the
kernels by assignment in gcn.py (or by deriving GCNModel from GCNFeaturePooling and deleting its lines 84-164):

    from modules.pooling import GCNFeaturePooling as P
    GCNModel.get_bound_of_images = staticmethod(P.get_bound_of_images)
    GCNModel.perceptual_feature_pooling = staticmethod(P.perceptual_feature_pooling)
"""
from .feature_pooling import GCNFeaturePooling

__all__ = ["GCNFeaturePooling"]
