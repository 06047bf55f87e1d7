"""`import segearth_segmentor` shim for the reference's eval.py:5 / demo.py:4."""
from clip_decontamination_b200.segearth_segmentor import Segmentor  # noqa: F401
