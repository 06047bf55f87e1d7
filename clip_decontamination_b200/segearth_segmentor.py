"""``Segmentor`` -- drop-in for the reference's upstream-style segmentor (segearth_segmentor.py:22-373,
used by demo.py:28-42): same kernels as ``SegmentorEx`` with the proposed extras off and the
``cls_token_lambda`` logit bias (segearth_segmentor.py:115-117,193)."""
import torch

from .compat import MODELS
from .segmentor import SegmentorEx


@MODELS.register_module()
class Segmentor(SegmentorEx):
    def __init__(self,
                 clip_type,
                 vit_type,
                 model_type,
                 name_path,
                 device=torch.device('cuda'),
                 ignore_residual=True,
                 prob_thd=0.0,
                 logit_scale=50,
                 slide_stride=112,
                 slide_crop=224,
                 cls_token_lambda=0,
                 bg_idx=0,
                 apply_sim_feat_up=True,
                 sim_feat_up_cfg=dict(model_name='jbu_one', model_path='your/model/path'),
                 **extensions):
        super().__init__(clip_type=clip_type, vit_type=vit_type, model_type=model_type, name_path=name_path,
                         device=device, ignore_residual=ignore_residual, prob_thd=prob_thd,
                         logit_scale=logit_scale, slide_stride=slide_stride, slide_crop=slide_crop,
                         cls_token_lambda=cls_token_lambda, bg_idx=bg_idx, apply_sim_feat_up=apply_sim_feat_up,
                         sim_feat_up_cfg=sim_feat_up_cfg, **extensions)
        self.output_cls_token = cls_token_lambda != 0
