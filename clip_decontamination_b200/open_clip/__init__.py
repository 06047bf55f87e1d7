"""Drop-in subset of the reference's vendored ``open_clip`` used by the segmentor:
``create_model``, ``tokenizer.tokenize`` (open_clip/__init__.py of the reference exports many more
training-time symbols that the segmentation path never reaches)."""
from . import tokenizer
from .factory import create_model
from .model import CLIP
from .model_configs import get_model_config, list_models
from .tokenizer import tokenize

__all__ = ['create_model', 'tokenizer', 'tokenize', 'CLIP', 'get_model_config', 'list_models']
