"""``SegmentorEx`` -- drop-in for the reference's mmseg segmentor (segmentor.py:25-622).

Same registry name, constructor kwargs, methods (``predict``, ``forward_slide``, ``forward_feature``,
``postprocess_result``, ``compute_padsize``) and attributes (``net``, ``query_features``, ``query_idx``,
``num_queries``, ``num_classes``, ``data_preprocessor``) as the reference, so the reference's ``eval.py``
/ ``cfg_*.py`` / ``cls_*.txt`` work unchanged when this module is importable as ``segmentor``
(see INTEGRATION.md).  The arithmetic runs in libclipseg (sm_100a); nothing here falls back to the CPU.

Differences that are visible to a caller (all documented in INTEGRATION.md):
* compute is bf16 (tcgen05) with fp32 residual stream / softmax / logits where the reference is fp16;
  ``precision='fp32'`` selects the CUDA-core verification mode;
* all crops of an image are processed as one batch;
* ``seg_logits`` (class probabilities) is only attached to the data samples when
  ``output_seg_logits=True`` -- the fused accumulate->argmax kernel does not write them otherwise.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from .compat import BaseSegmentor, SegDataPreProcessor, PixelData, MODELS
from .engine import SegEngine, compute_padsize as _compute_padsize, slide_windows
from .open_clip import create_model, tokenizer
from .outlier_suppression import OutlierSuppressionModule
from .prompts.imagenet_template import openai_imagenet_template
from .similarity_enhancement import SimilarityEnhancementModule
from .simfeatup_dev.upsamplers import get_upsampler

_CLIP_TABLE = {       # segmentor.py:69-112 (clip_type, 'B'/'L'/'H' in vit_type) -> (model name, pretrained)
    ('CLIP', 'B'): ('ViT-B/16', 'openai'), ('CLIP', 'L'): ('ViT-L-14', 'openai'),
    ('RemoteCLIP', 'B'): ('ViT-B/32', 'checkpoint/RemoteCLIP-ViT-B-32.pt'),
    ('RemoteCLIP', 'L'): ('ViT-L-14', 'checkpoint/RemoteCLIP-ViT-L-14.pt'),
    ('GeoRSCLIP', 'B'): ('ViT-B/32', 'checkpoint/RS5M_ViT-B-32.pt'),
    ('GeoRSCLIP', 'L'): ('ViT-L-14', 'checkpoint/RS5M_ViT-L-14.pt'),
    ('GeoRSCLIP', 'H'): ('ViT-H-14', 'checkpoint/RS5M_ViT-H-14.pt'),
    ('SkyCLIP', 'B'): ('ViT-B/32', 'checkpoint/SkyCLIP_ViT_B32_top50pct/epoch_20.pt'),
    ('SkyCLIP', 'L'): ('ViT-L-14', 'checkpoint/SkyCLIP_ViT_L14_top30pct_filtered_by_CLIP_laion_RS/epoch_20.pt'),
    ('OpenCLIP', 'B'): ('ViT-B/16', 'laion2b_s34b_b88k'), ('OpenCLIP', 'L'): ('ViT-L-14', 'laion2b_s32b_b82k'),
    ('MetaCLIP', 'B'): ('ViT-B-16-quickgelu', 'metaclip_fullcc'),
    ('MetaCLIP', 'L'): ('ViT-L/14-quickgelu', 'metaclip_fullcc'),
    ('ALIP', 'B'): ('ViT-B/32', 'checkpoint/ALIP_YFCC15M_B32.pt'),
}


def get_cls_idx(path):
    """segmentor.py:611-622 -- one class per line, ',' separates synonyms, only the newline is stripped."""
    with open(path, 'r') as f:
        name_sets = f.readlines()
    class_names, class_indices = [], []
    for idx, line in enumerate(name_sets):
        names_i = line.split(',')
        class_names += names_i
        class_indices += [idx for _ in range(len(names_i))]
    class_names = [item.replace('\n', '') for item in class_names]
    return class_names, class_indices


def _vit_key(vit_type: str) -> str:
    for k in ('B', 'L', 'H'):
        if k in vit_type:
            return k
    raise ValueError(f'cannot parse vit_type {vit_type!r}')


@MODELS.register_module()
class SegmentorEx(BaseSegmentor):
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
                 cls_token_lambda=0.0,
                 global_debias_factor=0.0,
                 bg_idx=0,
                 apply_sim_feat_up=False,
                 sim_feat_up_cfg=dict(model_name='jbu_one', model_path='your/model/path'),
                 apply_ctd=False,
                 apply_outlier_suppression=False,
                 outlier_suppression_cfg=None,
                 apply_self_attn_enhancement=False,
                 self_attn_enhancement_cfg=None,
                 apply_layer_fusion=False,
                 layer_fusion_lambda=0.5,
                 layer_fusion_threshold=0.7,
                 apply_similarity_enhancement=False,
                 similarity_enhancement_cfg=None,
                 result_dir=None,
                 heatmap_dir=None,
                 # ---- extensions (not in the reference) ----
                 precision='bf16',
                 output_seg_logits=False,
                 query_features=None,
                 net=None,
                 upsampler_state_dict=None,
                 use_graph=True,
                 ):
        data_preprocessor = SegDataPreProcessor(
            mean=[122.771, 116.746, 104.094],
            std=[68.501, 66.632, 70.323],
            bgr_to_rgb=True)
        super().__init__(data_preprocessor=data_preprocessor)
        # switches that are outside the CUDA hot path: accepted when off, refused when on
        for flag, name in ((apply_ctd, 'apply_ctd'), (apply_self_attn_enhancement, 'apply_self_attn_enhancement'),
                           (apply_layer_fusion, 'apply_layer_fusion')):
            if flag:
                raise NotImplementedError(f'{name}=True is not part of the B200 hot path (default off in the '
                                          f'reference, enabled by no config)')
        if clip_type == 'BLIP' or model_type == 'GEM':
            raise NotImplementedError('BLIP / GEM backbones are not part of the B200 hot path')
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('SegmentorEx runs on CUDA only (there is no CPU fallback)')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        if net is None:
            key = (clip_type, _vit_key(vit_type))
            if key not in _CLIP_TABLE:
                raise ValueError(f'unknown clip_type / vit_type {clip_type!r} / {vit_type!r}')
            name, pretrained = _CLIP_TABLE[key]
            net = create_model(name, pretrained=pretrained, precision='fp32' if precision == 'fp32' else 'fp16')
        self.net = net
        self.net.precision = 'fp32' if precision == 'fp32' else 'bf16'
        self.net.eval().to(device)
        self.tokenizer = tokenizer.tokenize
        self.clip_type, self.vit_type, self.model_type = clip_type, vit_type, model_type
        self.apply_sim_feat_up = apply_sim_feat_up
        self.cls_token_lambda = cls_token_lambda
        self.global_debias_factor = global_debias_factor
        self.bg_idx = bg_idx
        self.patch_size = self.net.visual.patch_size

        query_words, query_idx = get_cls_idx(name_path)
        self.num_queries = len(query_words)
        self.num_classes = max(query_idx) + 1
        self.query_idx = torch.Tensor(query_idx).to(torch.int64).to(device)
        if query_features is None:                                   # segmentor.py:157-174 (init-time)
            query_features = self._text_cache(query_words, device)
        self.query_features = query_features.to(device).float()
        assert self.query_features.shape[0] == self.num_queries
        self.dtype = self.query_features.dtype
        self.ignore_residual = ignore_residual
        self.logit_scale, self.prob_thd = logit_scale, prob_thd
        self.slide_stride, self.slide_crop = slide_stride, slide_crop
        self.apply_ctd = apply_ctd
        self.apply_layer_fusion = apply_layer_fusion
        self.layer_fusion_lambda, self.layer_fusion_threshold = layer_fusion_lambda, layer_fusion_threshold
        self.apply_self_attn_enhancement = apply_self_attn_enhancement

        self.apply_similarity_enhancement = apply_similarity_enhancement
        sim_cfg = None
        if apply_similarity_enhancement:                             # segmentor.py:196-220
            sim_cfg = dict(similarity_weight=1.0, temperature=1.0, add_self_similarity=True)
            if similarity_enhancement_cfg:
                sim_cfg.update(similarity_enhancement_cfg)
            self.net.visual.similarity_enhancer = SimilarityEnhancementModule(**sim_cfg)
        self.apply_outlier_suppression = apply_outlier_suppression
        out_cfg = None
        if apply_outlier_suppression:                                # segmentor.py:252-274
            out_cfg = dict(top_k=10)
            if outlier_suppression_cfg:
                out_cfg.update(outlier_suppression_cfg)
            self.net.visual.outlier_suppressor = OutlierSuppressionModule(top_k=out_cfg['top_k'])
            out_cfg['contamination_temp'] = self.net.visual.outlier_suppressor.contamination_temp
        self.result_dir, self.heatmap_dir = result_dir, heatmap_dir
        self.output_seg_logits = output_seg_logits or bool(heatmap_dir)
        self.use_graph = use_graph
        self.last_labels = None          # uint8 label maps behind the last predict / test_step (native evaluators)

        up_engine = None
        if self.apply_sim_feat_up:                                   # segmentor.py:278-284
            self.feat_dim = self.query_features.shape[-1]
            self.upsampler = get_upsampler(sim_feat_up_cfg['model_name'], self.feat_dim)
            if upsampler_state_dict is None:
                ckpt = torch.load(sim_feat_up_cfg['model_path'], map_location='cpu', weights_only=False)['state_dict']
                upsampler_state_dict = {k[10:]: v for k, v in ckpt.items()}
            self.upsampler.load_state_dict(upsampler_state_dict, strict=True)
            self.upsampler.precision = self.net.precision
            self.upsampler.to(device)
            up_engine = self.upsampler.engine(device, self.net.precision)
        self._device = device
        self.engine = SegEngine(self.net.visual_engine(device), self.query_features, query_idx,
                                model_type=model_type, ignore_residual=ignore_residual, prob_thd=prob_thd,
                                logit_scale=logit_scale, slide_stride=slide_stride, slide_crop=slide_crop,
                                cls_token_lambda=cls_token_lambda, global_debias_factor=global_debias_factor,
                                bg_idx=bg_idx, upsampler=up_engine, sim_cfg=sim_cfg, outlier_cfg=out_cfg)

    # ---- A13 / N3: prompt-ensembled class embeddings, cached on disk ------------------------------
    def _text_cache(self, query_words, device):
        """segmentor.py:157-174: per query word the 80 imagenet templates -> tokenizer -> text tower (TextEngine, the
        same tcgen05 GEMM / tensor-core attention kernels as the ViT) -> unit norm -> mean -> unit norm.  The result
        only depends on the text-side weights, the words and the templates, so it is cached under
        $CLIPSEG_CACHE_DIR (default ~/.cache/clip_decontamination_b200) keyed by a fingerprint of those."""
        import hashlib
        n_t = len(openai_imagenet_template)
        hsh = hashlib.sha1()
        hsh.update(repr((list(query_words), [t('X') for t in openai_imagenet_template], self.net.precision,
                         self.net.quick_gelu)).encode())
        for k, p in sorted(self.net.state_dict().items()):
            if k.startswith('visual.'):
                continue
            pf = p.detach().float().flatten()
            step = max(1, pf.numel() // 4096)
            hsh.update(k.encode())
            hsh.update(np.asarray(list(p.shape) + [float(pf.double().sum()), float(pf.double().abs().sum())]).tobytes())
            hsh.update(pf[::step].cpu().numpy().tobytes())
        cdir = os.environ.get('CLIPSEG_CACHE_DIR', os.path.join(os.path.expanduser('~'), '.cache', 'clip_decontamination_b200'))
        path = os.path.join(cdir, f'text_{hsh.hexdigest()}.npy')
        if os.path.isfile(path):
            try:
                qf = torch.from_numpy(np.load(path))
                if qf.shape[0] == len(query_words):
                    return qf
            except Exception:
                pass
        prompts = [temp(qw) for qw in query_words for temp in openai_imagenet_template]
        tokens = self.tokenizer(prompts)
        feats = []
        with torch.no_grad():
            for c0 in range(0, len(prompts), 4 * n_t):                # four classes (320 prompts, M = 24 640 rows) per pass
                feats.append(self.net.encode_text(tokens[c0:c0 + 4 * n_t].to(device)))
            f = torch.cat(feats).float().view(len(query_words), n_t, -1)
            f = f / f.norm(dim=-1, keepdim=True)
            f = f.mean(dim=1)
            qf = f / f.norm(dim=-1, keepdim=True)
        try:
            os.makedirs(cdir, exist_ok=True)
            tmp = path + f'.{os.getpid()}.tmp'
            with open(tmp, 'wb') as fh:
                np.save(fh, qf.cpu().numpy())
            os.replace(tmp, path)
        except OSError:
            pass                                                     # read-only home: recompute next time
        return qf

    # ------------------------------------------------------------------------------------------
    def _img3(self, img):
        if type(img) == list:
            img = img[0]
        if img.dim() == 4:
            if img.shape[0] != 1:
                raise ValueError('forward_feature / forward_slide take one image (the reference views with batch 1, '
                                 'segmentor.py:369,372); use predict() for batches')
            img = img[0]
        return img.to(self._device, torch.float32).contiguous()

    def forward_feature(self, img, logit_size=None, tile_h_idx=None, tile_w_idx=None):
        """Cosine logits of ONE crop [1,Q,h,w] resized to `logit_size` / the crop size (segmentor.py:286-392)."""
        img3 = self._img3(img)
        _, H, W = img3.shape
        saved = self.engine.crop
        self.engine.crop = 0
        try:
            logits, g = self.engine.crop_logits(img3)
        finally:
            self.engine.crop = saved
        size = (H, W) if logit_size is None else tuple(logit_size)
        return nn.functional.interpolate(logits, size=size, mode='bilinear')

    def forward_slide(self, img, img_metas, stride=112, crop_size=224):
        """Averaged cosine logits [1,Q,H0,W0] (segmentor.py:394-451)."""
        img3 = self._img3(img)
        eng = self.engine
        saved = eng.stride, eng.crop
        stride = stride[0] if isinstance(stride, (tuple, list)) else stride
        crop_size = crop_size[0] if isinstance(crop_size, (tuple, list)) else crop_size
        eng.stride, eng.crop = stride, crop_size
        try:
            _, _, avg = eng.segment(img3, None, want_logits=True)
        finally:
            eng.stride, eng.crop = saved
        logits = avg.unsqueeze(0)
        img_size = tuple(img_metas[0]['ori_shape'][:2])
        if img_size != tuple(logits.shape[-2:]):
            logits = nn.functional.interpolate(logits, size=img_size, mode='bilinear')
        return logits

    # ---- the fast path shared by predict / test_step / predict_u8 ---------------------------------
    def _labels(self, x, kind, ori_shapes):
        """x: B equally sized images laid out as `kind` (engine.graph), one tensor or a list of per-image tensors ->
        uint8 label maps on the device ([B,oh,ow] tensor, or a list when the images are resized to different
        ori_shapes).  Images that keep their size go through ONE CUDA-graph replay as a stacked batch; images that are
        resized to ori_shape (Resize in the test pipeline) are replayed one by one."""
        xs = list(x) if isinstance(x, (list, tuple)) else [x[i] for i in range(x.shape[0])]
        H, W = (xs[0].shape[0], xs[0].shape[1]) if kind == 'u8hwc' else (xs[0].shape[1], xs[0].shape[2])
        oris = [tuple(int(v) for v in o[:2]) for o in ori_shapes]
        if all(o == (H, W) for o in oris):
            self.last_labels = self.engine.segment_batch(x, kind, None, use_graph=self.use_graph)
        else:
            self.last_labels = [self.engine.segment_batch(xs[i].unsqueeze(0), kind, oris[i], use_graph=self.use_graph)[0]
                                for i in range(len(xs))]
        return self.last_labels

    def _attach(self, data_samples, labels, probs_list=None):
        out = []
        for i, lab in enumerate(labels):
            seg_pred = lab.to(torch.int64).unsqueeze(0)
            if data_samples is None:
                out.append(seg_pred)
                continue
            d = {'pred_sem_seg': PixelData(**{'data': seg_pred})}
            pr = probs_list[i] if probs_list is not None else None
            if pr is not None:
                d['seg_logits'] = PixelData(**{'data': pr})
            data_samples[i].set_data(d)
            if self.result_dir or self.heatmap_dir:
                self._dump(i, data_samples[i], lab, pr)
        if data_samples is None:
            return out[0] if len(out) == 1 else torch.stack(out)   # reference returns image 0 (batch 1)
        return data_samples

    @torch.no_grad()
    def predict(self, inputs, data_samples):
        """segmentor.py:453-473.  inputs [B,3,H,W] normalised floats (any float dtype, any device).  All B images go
        through the engine as ONE batch (B x n_crops crops per kernel launch, CUDA-graph replay per input shape)."""
        if data_samples is not None:
            batch_img_metas = [ds.metainfo for ds in data_samples]
        else:
            batch_img_metas = [dict(ori_shape=inputs.shape[2:], img_shape=inputs.shape[2:],
                                    pad_shape=inputs.shape[2:], padding_size=[0, 0, 0, 0])] * inputs.shape[0]
        oris = [m['ori_shape'] for m in batch_img_metas]
        if self.output_seg_logits:                                   # probabilities requested: eager, per image
            labels, probs = [], []
            for i in range(inputs.shape[0]):
                img = inputs[i].to(self._device, torch.float32).contiguous()
                lab, pr, _ = self.engine.segment(img, tuple(oris[i][:2]), want_probs=True)
                labels.append(lab)
                probs.append(pr)
            return self._attach(data_samples, labels, probs)
        x = inputs if inputs.dtype == torch.float32 else inputs.float()
        return self._attach(data_samples, self._labels(x.contiguous(), 'f32', oris))

    @torch.no_grad()
    def test_step(self, data):
        """mmengine BaseModel.test_step (what Runner.test() calls, eval.py:86-87): data = dict(inputs=[uint8 CHW BGR
        tensors], data_samples=[SegDataSample]).  The SegDataPreProcessor arithmetic (segmentor.py:64-67) is fused
        into the kernels that read the image, so the raw bytes are uploaded as they are; anything that is not a
        list of equally sized uint8 CHW images takes the generic data_preprocessor + predict route."""
        inputs, samples = data['inputs'], data.get('data_samples')
        fast = (isinstance(inputs, (list, tuple)) and len(inputs) > 0 and not self.output_seg_logits
                and all(torch.is_tensor(t) and t.dtype == torch.uint8 and t.dim() == 3 and t.shape[0] == 3
                        and t.shape == inputs[0].shape for t in inputs))
        if not fast:
            data = self.data_preprocessor(data, False)
            return self.predict(data['inputs'], data['data_samples'])
        shape = tuple(inputs[0].shape[1:])
        if samples is not None:
            oris = [ds.metainfo.get('ori_shape', shape) for ds in samples]
        else:
            oris = [shape] * len(inputs)
        return self._attach(samples, self._labels(list(inputs), 'u8chw', oris))

    @torch.no_grad()
    def predict_u8(self, img_hwc_bgr_u8, labels_out=None, use_graph=True, copy_out=True):
        """uint8 HWC BGR image [H,W,3] or batch [B,H,W,3] (pinned host or device, as cv2.imread returns it) -> uint8
        labels on the device.  Equivalent to data_preprocessor + predict; the launch sequence is replayed from a CUDA
        graph captured on first use of each input shape.  copy_out=False returns the graph's static output buffer
        (valid until the next call)."""
        return self.engine.segment_u8(img_hwc_bgr_u8, labels_out, use_graph, copy_out)

    def postprocess_result(self, seg_logits, data_samples):
        """segmentor.py:475-499 on given averaged logits [B,Q,H,W] (runs the fused kernel with one
        full-size window per image)."""
        from . import ops
        B, Q, H, W = seg_logits.shape
        out = []
        for i in range(B):
            lg = seg_logits[i].to(self._device, torch.float32).contiguous().unsqueeze(0)
            win = torch.tensor([(0, 0, H, W)], dtype=torch.int32, device=self._device)
            labels = torch.empty((H, W), dtype=torch.uint8, device=self._device)
            probs = torch.empty((self.num_classes, H, W), dtype=torch.float32, device=self._device)
            ops.accum_argmax(lg, win, H, W, 0, 0, H, W, H, W, self.engine.query_idx, self.num_classes,
                             float(self.logit_scale), float(self.prob_thd), int(self.bg_idx), labels, probs)
            seg_pred = labels.to(torch.int64).unsqueeze(0)
            if data_samples is None:
                return seg_pred
            data_samples[i].set_data({'seg_logits': PixelData(**{'data': probs}),
                                      'pred_sem_seg': PixelData(**{'data': seg_pred})})
        return data_samples

    def compute_padsize(self, H: int, W: int, patch_size: int):
        return _compute_padsize(H, W, patch_size)

    # ---- optional PNG dumps (segmentor.py:501-531,568-608): colourised on the GPU, one D2H copy + cv2.imwrite -----
    def _palette_bgr(self):
        """_generate_palette (segmentor.py:568-579), rows flipped to BGR for cv2.imwrite (:516)."""
        if getattr(self, '_pal', None) is None:
            import colorsys
            n = int(self.num_classes)
            pal = []
            for idx in range(n):
                r, g, b = colorsys.hsv_to_rgb((idx / max(1, n)) % 1.0, 0.75, 1.0 if idx != self.bg_idx else 0.2)
                pal.append([int(b * 255), int(g * 255), int(r * 255)])
            self._pal = torch.tensor(pal, dtype=torch.uint8, device=self._device)
        return self._pal

    def _jet_bgr(self):
        """cv2.COLORMAP_JET as a 256-entry BGR table (segmentor.py:601-603: the RGB conversion there is undone by the
        [:, :, ::-1] at :527, so the file holds cv2's own BGR colours)."""
        if getattr(self, '_jet', None) is None:
            import cv2
            lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
            self._jet = torch.from_numpy(np.ascontiguousarray(lut)).to(self._device)
        return self._jet

    def _dump(self, i, sample, labels_u8, probs):
        import cv2
        from . import ops
        meta = getattr(sample, 'metainfo', {}) or {}
        stem = next((os.path.splitext(os.path.basename(meta[k]))[0] for k in
                     ('img_path', 'ori_path', 'filename', 'ori_filename') if meta.get(k)), f'sample_{i}')
        if self.result_dir:
            os.makedirs(self.result_dir, exist_ok=True)
            img = ops.colorize(labels_u8.contiguous(), self._palette_bgr())
            cv2.imwrite(os.path.join(self.result_dir, f'{stem}.png'), img.cpu().numpy())
        if self.heatmap_dir and probs is not None:
            os.makedirs(self.heatmap_dir, exist_ok=True)
            img = ops.heatmap(probs.contiguous(), self._jet_bgr())
            cv2.imwrite(os.path.join(self.heatmap_dir, f'{stem}.png'), img.cpu().numpy())

    # mmseg abstract methods (unused, as in the reference segmentor.py:548-566)
    def _forward(self, *a, **k):
        pass

    def inference(self, img, batch_img_metas):
        pass

    def encode_decode(self, inputs, batch_img_metas):
        pass

    def extract_feat(self, inputs):
        pass

    def loss(self, inputs, data_samples):
        pass

    # aliases named in BASELINE.json's north_star
    slide_inference = forward_slide


SegEarthSegmentation = SegmentorEx
