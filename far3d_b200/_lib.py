"""ctypes binding of libfar3d_sm100.so (the C ABI declared in include/far3d_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FAR3D_LIB_PATH: an alternative build of the same library (A/B runs of kernel variants from tools/; still no fallback)
LIB_PATH = os.environ.get('FAR3D_LIB_PATH') or os.path.join(_HERE, 'lib', 'libfar3d_sm100.so')

c_int, c_i64, c_f, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/far3d_b200.h
SIGNATURES = {
    'far3d_last_error': [],
    'far3d_abi_version': [],
    'far3d_launch_count': [],
    'far3d_add_launches': [c_i64],
    'far3d_deform_agg_fwd': [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_vp] + [c_int] * 8 + [c_vp],
    'far3d_deform_agg_debug': [c_vp, c_vp, c_vp, c_f, c_f, c_vp, c_vp, c_vp] + [c_int] * 5 + [c_vp],
    'far3d_msda_fwd': [c_vp] * 6 + [c_int] * 7 + [c_vp],
    'far3d_dfa_weights_softmax': [c_vp] * 3 + [c_int] * 5 + [c_vp],
    'far3d_cam_logits': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f, c_vp, c_vp],
    'far3d_dfa_prepare_supported': [c_int] * 5,
    'far3d_dfa_prepare': [c_vp] * 6 + [c_f, c_f] + [c_int] * 8 + [c_vp] * 5,
    'far3d_deform_agg_gather': [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp] + [c_int] * 8 + [c_vp],
    'far3d_linear_mma': [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    'far3d_linear_f32': [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    'far3d_layernorm': [c_vp] * 5 + [c_int, c_int, c_f, c_int, c_int, c_vp],
    'far3d_mha_fwd': [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int] + [c_int] * 5 + [c_vp],
    'far3d_mha_fwd_masked': [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int] + [c_int] * 5 + [c_vp, c_vp],
    'far3d_roi_select': [c_vp, c_vp, c_vp, c_vp] + [c_int] * 5 + [c_f, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    'far3d_query2d_lift_workspace_ints': [c_int],
    'far3d_query2d_lift': [c_vp] * 4 + [c_int] * 3 + [c_vp] + [c_int] * 7 + [c_f] * 3 + [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    'far3d_ctx_gather': [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    'far3d_box_decode': [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_f, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    'far3d_memory_post_update': [c_vp] * 3 + [c_int] * 6 + [c_vp] * 13 + [c_vp],
    'far3d_memory_pre_update': [c_int] * 3 + [c_vp] * 14 + [c_vp],
    'far3d_pos2posemb3d': [c_vp, c_vp, c_int, c_int, c_vp],
    'far3d_pos2posemb1d': [c_vp, c_int, c_vp, c_int, c_int, c_vp],
    'far3d_nerf_posenc': [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'far3d_mln_flatten': [c_vp] * 4 + [c_int] * 6 + [c_vp],
    'far3d_mln_tokens': [c_vp] * 4 + [c_int] * 3 + [c_vp],
    'far3d_conv2d_umma': [c_vp, c_vp] + [c_int] * 6 + [c_vp, c_vp, c_vp] + [c_int] * 4 + [c_vp, c_int, c_int, c_i64,
                          c_vp, c_vp, c_int, c_int, c_vp],
    'far3d_conv2d_umma_pool': [c_vp, c_vp] + [c_int] * 6 + [c_vp, c_vp, c_vp] + [c_int] * 2 + [c_vp, c_int, c_int, c_vp, c_vp, c_vp],
    'far3d_conv2d_umma_mx': [c_vp, c_vp] + [c_int] * 7 + [c_vp, c_vp, c_int, c_vp] + [c_int] * 4 + [c_vp, c_int, c_int, c_i64,
                             c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'far3d_conv2d_umma_pool_mx': [c_vp, c_vp] + [c_int] * 7 + [c_vp, c_vp, c_int, c_vp] + [c_int] * 2 + [c_vp, c_int, c_int, c_vp,
                                  c_vp, c_vp],
    'far3d_conv_pool_workspace_floats': [c_int] * 4,
    'far3d_conv2d_f32': [c_vp] + [c_int] * 6 + [c_vp, c_vp] + [c_int] * 4 + [c_vp, c_int, c_int, c_vp],
    'far3d_stem_conv': [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    'far3d_maxpool3x3s2': [c_vp, c_vp] + [c_int] * 7 + [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'far3d_global_avgpool': [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'far3d_ese_gate': [c_vp] * 4 + [c_int, c_int, c_vp],
    'far3d_ese_apply': [c_vp] * 5 + [c_int] * 5 + [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp],
    'far3d_upsample_add': [c_vp, c_vp] + [c_int] * 6 + [c_vp, c_vp, c_int, c_vp],
    'far3d_groupnorm_nhwc': [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    'far3d_split_fp16': [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp],
    'far3d_linear_umma': [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    'far3d_merge_fp16': [c_vp, c_vp, c_vp, c_i64, c_vp],
    'far3d_merge_fp16_strided': [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp],
    'far3d_split_planes': [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp],
    'far3d_normalize_u8': [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp],
    'far3d_resample_ksize': [c_int, c_int],
    'far3d_resample_coeffs': [c_int, c_int, c_vp, c_vp],
    'far3d_resize_crop_u8': [c_vp] + [c_int] * 4 + [c_vp, c_vp, c_int, c_vp, c_vp] + [c_int] * 8 + [c_vp, c_vp, c_int, c_vp],
    'far3d_deform_agg_tune': [c_int, c_int],
    'far3d_mha_tune': [c_int],
    'far3d_conv_umma_tune': [c_int, c_int],
    'far3d_conv_umma_tune2': [c_int, c_int],
    'far3d_conv_umma_debug': [c_vp],
    'far3d_conv_umma_tune4': [c_int],
    'far3d_conv_umma_tune6': [c_f],
    'far3d_conv_umma_tune7': [c_int],
    'far3d_conv_umma_tune8': [c_int],
}
_RESTYPE = {'far3d_last_error': ctypes.c_char_p, 'far3d_query2d_lift_workspace_ints': c_i64, 'far3d_conv_pool_workspace_floats': c_i64, 'far3d_launch_count': c_i64, 'far3d_add_launches': None, 'far3d_deform_agg_tune': None, 'far3d_mha_tune': None, 'far3d_conv_umma_tune': None,
            'far3d_conv_umma_tune2': None, 'far3d_conv_umma_debug': None, 'far3d_conv_umma_tune4': None, 'far3d_conv_umma_tune6': None, 'far3d_conv_umma_tune7': None, 'far3d_conv_umma_tune8': None}

_lib = None


class Far3DNativeError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it was not built: the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Far3DNativeError(
            f'{LIB_PATH} not found - build it with `python -m far3d_b200.build` (or __graft_entry__.build()); '
            'far3d_b200 has no CPU / PyTorch fallback')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so does not export what the header declares
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise Far3DNativeError(f'{name} failed ({rc}): {lib.far3d_last_error().decode()}')


def launch_count():
    return int(load().far3d_launch_count())
