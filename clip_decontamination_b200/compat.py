"""mmseg / mmengine compatibility.  When the real packages are installed they are used unchanged
(registry, BaseSegmentor, SegDataPreProcessor, PixelData), which is what makes the segmentor classes
drop into the reference's eval.py.  They are NOT installed in the build image, so minimal stand-ins with
the same call surface are provided for the standalone API, the tests and the benchmark."""
import torch
import torch.nn as nn

try:                                                        # pragma: no cover - not installed here
    from mmseg.models.segmentors import BaseSegmentor
    from mmseg.models.data_preprocessor import SegDataPreProcessor
    from mmseg.registry import MODELS
    from mmengine.structures import PixelData
    HAVE_MMSEG = True
except Exception:
    HAVE_MMSEG = False

    class _Registry:
        def __init__(self):
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                self.module_dict[name or cls.__name__] = cls
                return cls
            return deco(module) if module is not None else deco

        def build(self, cfg: dict):
            cfg = dict(cfg)
            return self.module_dict[cfg.pop('type')](**cfg)

    MODELS = _Registry()

    class PixelData:
        def __init__(self, data=None, **kw):
            self.data = data

    class SegDataSample:
        def __init__(self, metainfo=None):
            self.metainfo = dict(metainfo or {})

        def set_data(self, d: dict):
            for k, v in d.items():
                setattr(self, k, v)

    class SegDataPreProcessor(nn.Module):
        """(x[::-1] - mean) / std on the model device (mmseg 1.2.2 semantics, batch of CHW uint8 BGR)."""

        def __init__(self, mean=None, std=None, bgr_to_rgb=False, **kw):
            super().__init__()
            self.bgr_to_rgb = bgr_to_rgb
            self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1), False)
            self.register_buffer('std', torch.tensor(std, dtype=torch.float32).view(-1, 1, 1), False)

        def forward(self, data: dict, training: bool = False):
            inputs = []
            for x in data['inputs']:
                x = x.to(self.mean.device)
                if self.bgr_to_rgb:
                    x = x[[2, 1, 0]]
                inputs.append((x.float() - self.mean) / self.std)
            return dict(inputs=torch.stack(inputs), data_samples=data.get('data_samples'))

    class BaseSegmentor(nn.Module):
        def __init__(self, data_preprocessor=None, init_cfg=None):
            super().__init__()
            self.data_preprocessor = data_preprocessor

        def forward(self, inputs, data_samples=None, mode='predict'):
            if mode == 'predict':
                return self.predict(inputs, data_samples)
            raise NotImplementedError(mode)

        def test_step(self, data):
            data = self.data_preprocessor(data, False)
            return self.forward(data['inputs'], data['data_samples'], mode='predict')
