"""`import segmentor` shim for the reference's eval.py:4 -- put this directory first on PYTHONPATH."""
from clip_decontamination_b200.segmentor import *          # noqa: F401,F403
from clip_decontamination_b200.segmentor import SegmentorEx, SegEarthSegmentation, get_cls_idx  # noqa: F401
