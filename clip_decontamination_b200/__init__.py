"""clip_decontamination_b200 -- B200-native (sm_100a) implementation of the dense CLIP open-vocabulary
segmentation hot path of UserNameUnavailableIsUnavailable/CLIP-Decontamination.

Public surface (mirrors the reference's module names):
  segmentor.SegmentorEx / SegEarthSegmentation, segearth_segmentor.Segmentor,
  open_clip.create_model / tokenizer, simfeatup_dev.upsamplers.get_upsampler,
  engine.{VisualEngine, JBUEngine, SegEngine}, ops (C-ABI wrappers), dist (multi-GPU evaluation).
Importing the package does not load CUDA; the kernels live in libclipseg.so (see include/clipseg.h).
"""
__version__ = '0.1.0'
