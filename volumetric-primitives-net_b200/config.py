"""Configuration constants the hot-path drop-ins read, with the reference's defaults (config.py:1-59).

The reference binds several of these at import time as default arguments (chamfer_distance.py:3,10;
silhouette.py:5,11; vp_diverse.py:4); the drop-ins do the same.  Only the constants the primitive
assembly + loss path consumes are mirrored here; a caller that keeps its own `config` module first on
sys.path (the reference's layout) overrides this file.
"""
DEVICE = 'cuda'

SAMPLE_NUM = 128
BATCH_SIZE = 8
CD_W1 = 1.0
CD_W2 = 1.0
L_VIEW_CD = 1.0
L_CAN_CD = 0.0
L_SIL = 0.0
L_VP_DIV = 0.1
L_EMD = 1.0

MANUAL_SEED = 1234
SILHOUETTE_LOSS_FUNC = 'L1'  # L1 or MSE

CUBOID_NUM = 0
SPHERE_NUM = 16
CONE_NUM = 0
VP_NUM = CUBOID_NUM + SPHERE_NUM + CONE_NUM

IMG_SIZE = 128
IS_VIEW_CENTER = True
