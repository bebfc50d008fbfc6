"""Hot-path constants of lib/model/utils/config.py (values cited by line there)."""
import numpy as np


class _Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


cfg = _Cfg(
    TRAIN=_Cfg(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_MIN_SIZE=8,  # :96-102
               BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0), BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),     # :77-78
               DIM_NORMALIZE_MEANS=(1.6, 1.5, 4.0, 0.0, 0.0), DIM_NORMALIZE_STDS=(0.5,) * 5,            # :81-82
               TRUNCATED=False),
    TEST=_Cfg(NMS=0.3, RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300,              # :124-132
              RPN_MIN_SIZE=16),
    RESNET=_Cfg(FIXED_BLOCKS=1),                                                                         # :154
    PIXEL_MEANS=np.array([[[102.9801, 115.9465, 122.7717]]]),                                            # :170
    KPTS_GRID=28, POOLING_SIZE=7, USE_GPU_NMS=True,                                                      # :173,204,196
    ANCHOR_RATIOS=[0.5, 1, 2], FEAT_STRIDE=[16, ],                                                       # :210,213
    FPN_ANCHOR_SCALES=[32, 64, 128, 256, 512], FPN_FEAT_STRIDES=[4, 8, 16, 32, 64], FPN_ANCHOR_STRIDE=1,  # :216-222
)
