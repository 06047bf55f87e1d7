"""``get_upsampler`` with the reference signature (simfeatup_dev/upsamplers.py:353-369).

Only the guided JBU upsamplers that the configs select are built ('jbu_one' -- every cfg_*.py through
base_config.py:22-24 -- and 'jbu_stack', the one with shipped checkpoints).  The modules hold parameters
under the reference's key names so ``load_state_dict(strict=True)`` of the Lightning checkpoints works
(segmentor.py:281-283); ``forward`` runs the CUDA ``JBUEngine``.
"""
import torch
import torch.nn as nn


class _JBULearnedRange(nn.Module):
    """parameter holder for JBULearnedRange (simfeatup_dev/upsamplers.py:202-228)"""

    def __init__(self, guidance_dim, feat_dim, key_dim, radius):
        super().__init__()
        self.radius = radius
        d2 = (2 * radius + 1) ** 2
        self.range_temp = nn.Parameter(torch.tensor(0.0))
        self.range_proj = nn.Sequential(nn.Conv2d(guidance_dim, key_dim, 1, 1), nn.GELU(), nn.Dropout2d(.1),
                                        nn.Conv2d(key_dim, key_dim, 1, 1))
        self.fixup_proj = nn.Sequential(nn.Conv2d(guidance_dim + d2, d2, 1, 1), nn.GELU(), nn.Dropout2d(.1),
                                        nn.Conv2d(d2, d2, 1, 1))
        self.sigma_spatial = nn.Parameter(torch.tensor(1.0))


class _JBUBase(nn.Module):
    name = ''

    def __init__(self, feat_dim):
        super().__init__()
        self.feat_dim = feat_dim
        self.fixup_proj = nn.Sequential(nn.Dropout2d(0.2), nn.Conv2d(feat_dim, feat_dim, kernel_size=1))
        self.precision = 'bf16'
        self._engine = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def half(self):            # segmentor.py:280 calls .cuda().half(); the CUDA path computes in bf16
        return self

    def engine(self, device, precision=None):
        from ..engine import JBUEngine
        precision = precision or self.precision
        if self._engine is None or self._engine.device != torch.device(device) or \
                self._engine.cdt != (torch.bfloat16 if precision == 'bf16' else torch.float32):
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._engine = JBUEngine(self.name, sd, self.feat_dim, precision, device)
        return self._engine

    @torch.no_grad()
    def forward(self, source, guidance):
        """source [B,C,h,w], guidance [B,3,H,W] (H=16h) -> [B,C,16h,16w]  (upsamplers.py:320-325)"""
        B, C, h, w = source.shape
        GB, _, H, W = guidance.shape
        assert B == GB                                                      # upsamplers.py:256
        if not source.is_cuda:
            raise RuntimeError('the JBU upsampler runs on CUDA only (no CPU fallback)')
        eng = self.engine(source.device)
        feats = source.permute(0, 2, 3, 1).reshape(B * h * w, C).to(eng.cdt).contiguous()
        tall = guidance.detach().float().permute(1, 0, 2, 3).reshape(3, B * H, W).contiguous()
        wins = torch.tensor([(b * H, 0, H, W) for b in range(B)], dtype=torch.int32, device=source.device)
        out = eng.upsample(feats, h, w, tall, wins, H, W)
        return out.view(B, 16 * h, 16 * w, C).permute(0, 3, 1, 2).to(source.dtype)


class JBUOne(_JBUBase):
    name = 'jbu_one'

    def __init__(self, feat_dim, *args, **kwargs):
        super().__init__(feat_dim)
        self.up = _JBULearnedRange(3, feat_dim, 32, radius=5)


class JBUStack(_JBUBase):
    name = 'jbu_stack'

    def __init__(self, feat_dim, *args, **kwargs):
        super().__init__(feat_dim)
        self.up1 = _JBULearnedRange(3, feat_dim, 32, radius=3)
        self.up2 = _JBULearnedRange(3, feat_dim, 32, radius=3)
        self.up3 = _JBULearnedRange(3, feat_dim, 32, radius=3)
        self.up4 = _JBULearnedRange(3, feat_dim, 32, radius=3)


_UNBUILT = ('bilinear', 'resize_conv', 'carafe', 'sapa', 'ifa')


def get_upsampler(upsampler, dim):
    if upsampler == 'jbu_stack':
        return JBUStack(dim)
    if upsampler == 'jbu_one':
        return JBUOne(dim)
    if upsampler in _UNBUILT:
        raise NotImplementedError(f"upsampler {upsampler!r} is never selected by a config of the reference and is "
                                  f"not part of the CUDA hot path")
    raise ValueError(f"Unknown upsampler {upsampler}")
