"""ctypes binding of libclipseg.so (the C ABI declared in include/clipseg.h).

There is no CPU fallback: importing this module fails loudly when the shared library has not
been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C clip_decontamination_b200/csrc``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libclipseg.so')

F32, BF16, F16, U8 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_QUICKGELU = 0, 1, 2
ATTN = dict(STD=0, Experimental=1, SCLIP=2, ClearCLIP=3, SFP=4, vanilla=5, SegEarth=6, MaskCLIP=7, CAUSAL=8)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: the CUDA extension has not been built. There is no CPU fallback; '
        f'run `make -C {os.path.join(_HERE, "csrc")}` (needs nvcc, targets sm_100a).')

lib = C.CDLL(LIB_PATH)

_p, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_f3 = C.c_float * 3


class CsegImage(C.Structure):
    """struct cseg_image of include/clipseg.h (host struct; `data` is a device pointer)."""
    _fields_ = [('data', C.c_void_p), ('dtype', C.c_int), ('H', C.c_int), ('W', C.c_int), ('img_h', C.c_int),
                ('stride_img', C.c_longlong), ('stride_c', C.c_longlong), ('stride_y', C.c_longlong),
                ('stride_x', C.c_longlong), ('chan', C.c_int * 3), ('mean', C.c_float * 3), ('std', C.c_float * 3)]


_img = C.POINTER(CsegImage)


class CsegJbuShare(C.Structure):
    """struct cseg_jbu_share of include/clipseg.h."""
    _fields_ = [('windows', C.c_void_p), ('shift', C.c_int), ('pitch', C.c_int)]


_shr = C.POINTER(CsegJbuShare)

SIGNATURES = {
    'cseg_version': (_i, []),
    'cseg_last_error': (_i, [C.c_char_p, C.c_size_t]),
    'cseg_launch_count': (_ll, []),
    'cseg_preprocess_u8': (_i, [_p, _i, _i, _f3, _f3, _p, _p]),
    'cseg_patchify': (_i, [_img, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    'cseg_gather_rows': (_i, [_p, _p, _p, _ll, _i, _i, _p, _p]),
    'cseg_embed_tokens': (_i, [_p, _p, _p, _i, _i, _i, _p, _p]),
    'cseg_embed_tokens_ln': (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _f, _p, _p]),
    'cseg_layernorm': (_i, [_p, _i, _i, _p, _p, _f, _i, _p, _p]),
    'cseg_gemm': (_i, [_i, _p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, _f, _i, _i, _p, _i, _p]),
    'cseg_gemm_blockdiag': (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    'cseg_gemm_reference': (_i, [_i, _p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, _f, _i, _i, _p, _i, _p]),
    'cseg_attention': (_i, [_i, _p, _i, _i, _i, _i, _i, _p, _f, _p, _p, _p]),
    'cseg_simmap': (_i, [_p, _i, _i, _i, _f, _i, _p, _p]),
    'cseg_simmap_tc': (_i, [_p, _i, _i, _i, _f, _p, _p, _i, _p]),
    'cseg_attention_experimental_tc': (_i, [_p, _i, _i, _i, _p, _f, _p, _p]),
    'cseg_outlier_suppress': (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _f, _p, _p, _p]),
    'cseg_cls_debias': (_i, [_p, _i, _i, _i, _f, _i, _p, _i, _i, _p, _p]),
    'cseg_jbu_guidance': (_i, [_img, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    'cseg_jbu_range_proj': (_i, [_p, _i, _i, _p, _p, _p, _p, _i, _p, _p]),
    'cseg_jbu_range_kernel': (_i, [_i, _p, _p, _i, _i, _i, _i, _i, _f, _f, _i, _p, _i, _i, _p]),
    'cseg_jbu_apply': (_i, [_i, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p]),
    'cseg_jbu_share_rows': (_i, [_i, _i, _i]),
    'cseg_jbu_range_kernel_border': (_i, [_p, _p, _shr, _i, _i, _i, _i, _f, _f, _p, _i, _i, _p]),
    'cseg_jbu_composite_image': (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    'cseg_jbu_apply_shared': (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _shr, _i, _i, _p, _p, _p]),
    'cseg_norm_sim': (_i, [_i, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    'cseg_fixup_norm_sim': (_i, [_i, _p, _i, _p, _i, _i, _i, _i, _p, _f, _p, _i, _p, _p, _p, _p]),
    'cseg_jbu_guidance_proj': (_i, [_img, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p]),
    'cseg_jbu_kernel_fixup': (_i, [_i, _p, _i, _p, _i, _p, _p, _i, _p, _i, _i, _p, _i, _p]),
    'cseg_basis_logits': (_i, [_i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p]),
    'cseg_accum_argmax': (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _i, _f, _f, _i,
                               _p, _p, _p, _p]),
    'cseg_colorize': (_i, [_p, _ll, _p, _i, _p, _p]),
    'cseg_heatmap': (_i, [_p, _i, _ll, _p, _p, _p]),
    'cseg_iou_hist': (_i, [_p, _p, _ll, _i, _i, _p, _p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _fn.restype, _fn.argtypes = _res, _args


class ClipSegError(RuntimeError):
    pass


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.cseg_last_error(buf, 512)
    return buf.value.decode(errors='replace')


def check(rc: int):
    if rc != 0:
        raise ClipSegError(f'libclipseg error {rc}: {last_error()}')


def launch_count() -> int:
    return int(lib.cseg_launch_count())
